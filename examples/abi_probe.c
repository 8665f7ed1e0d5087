/* Plain-C consumer of include/ft3d.h: proves the boundary is a C ABI (no C++ types, no torch) that links from C.
 * Calls only host-side entry points, so it also runs on a box without a GPU.
 *   gcc -std=c99 -Iinclude examples/abi_probe.c -Lfusiontransformer_b200 -lft3d -Wl,-rpath,$PWD/fusiontransformer_b200 */
#include <stdio.h>
#include "ft3d.h"

int main(void) {
  printf("ft3d version %d\n", ft3d_version());
  printf("table_capacity(1000) = %lld\n", (long long)ft3d_table_capacity(1000));
  printf("conv_packed_bytes(27, 96, 128) = %zu\n", ft3d_conv_packed_bytes(27, 96, 128));
  /* argument validation precedes every CUDA call: a bad shape is an error code + message, not a crash */
  int rc = ft3d_conv_pairs_tc((const void*)256, (const int32_t*)256, (const int32_t*)256, 27, 0, 1000, 20, 64,
                              (const void*)256, (float*)256, (ft3d_stream_t)0);
  printf("bad shape -> rc %d: %s\n", rc, ft3d_last_error());
  return rc == 0;
}
