/* Plain-C consumer of include/ultrasphere_b200.h: proves the boundary is a C ABI (no C++ or torch
 * types), compiles a plan for create_spherical() (`ba`) and exercises the argument checking that
 * needs no GPU.  Built and run by tests/test_host_cpu.py::test_header_is_plain_c_and_links. */
#include <stdio.h>
#include <string.h>

#include "ultrasphere_b200.h"

int main(void) {
  us_plan_t plan;
  /* pre-order program of `ba`: node 0 = b (cos leaf 2, sin internal), node 1 = a (leaves 0, 1) */
  const uint8_t kind[2] = {US_NODE_B, US_NODE_A};
  const int32_t slot[2] = {0, 1};
  const int32_t leaves[3] = {2, 0, 1};
  int rc;
  if (us_abi_version() != US_ABI_VERSION) return 1;
  if (!us_version() || strlen(us_version()) == 0) return 2;
  rc = us_plan_compile(2, kind, slot, leaves, &plan);
  if (rc != 0 || plan.n_nodes != 2 || plan.n_leaves != 3 || plan.max_stack != 0) return 3;
  /* an `a` that has nothing to resume, followed by another node: malformed */
  {
    const uint8_t bad[2] = {US_NODE_A, US_NODE_A};
    const int32_t bad_leaves[4] = {0, 1, 2, 3};
    if (us_plan_compile(2, bad, slot, bad_leaves, &plan) != US_EPLAN) return 4;
    if (strlen(us_last_error_string()) == 0) return 5;
  }
  rc = us_plan_compile(2, kind, slot, leaves, &plan);
  /* n = 0 is a no-op that must not touch the device */
  {
    const int64_t shape[1] = {0};
    const void* in[3] = {0, 0, 0};
    void* out[3] = {0, 0, 0};
    const uint32_t mask[3] = {1, 1, 1};
    if (us_to_cartesian(&plan, US_F64, 0, 1, shape, in, mask, 0, 0, out, 0) != 0) return 6;
    if (us_from_cartesian(&plan, US_F32, 0, 1, shape, in, mask, 0, out, 0) != 0) return 7;
    if (us_to_cartesian(&plan, 7, 0, 1, shape, in, mask, 0, 0, out, 0) != US_EINVAL) return 8;
  }
  printf("abi %d, library %s\n", us_abi_version(), us_version());
  return 0;
}
