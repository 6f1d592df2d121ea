"""CPU oracle of the ultrasphere hot path — TEST INFRASTRUCTURE, never imported by the product.

vks_oracle.py  NumPy restatement of the reference (pinned to tests/golden by gen_golden.py)
refload.py     loader of the UNMODIFIED reference from /root/reference (build container only)
gen_golden.py  writes tests/golden/* from the unmodified reference
stubs/         stand-ins for third-party packages the reference imports but this image lacks

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it.
"""
