// Bookkeeping for programmatic dependent launch (see td3d_common.cuh): which streams currently end in a kernel of this
// library, so that the next launch on them may start its grid before the previous one has drained.
#include "td3d_common.cuh"

#include <stdlib.h>

namespace td3d {

namespace {
struct PdlState {
  int enabled = -1;
  int n = 0;
  cudaStream_t chain[4];
};
thread_local PdlState g_pdl;
}  // namespace

bool pdl_take(cudaStream_t st) {
  PdlState& s = g_pdl;
  if (s.enabled < 0) {
    const char* e = getenv("TD3D_PDL");
    s.enabled = (e && atoi(e) == 0) ? 0 : 1;
  }
  if (!s.enabled) return false;
  for (int i = 0; i < s.n; ++i)
    if (s.chain[i] == st) return true;
  if (s.n == 4) s.n = 0;                       // more streams than slots: forget (costs one plain launch each)
  s.chain[s.n++] = st;
  return false;
}

void pdl_break(cudaStream_t st) {
  PdlState& s = g_pdl;
  for (int i = 0; i < s.n; ++i)
    if (s.chain[i] == st) { s.chain[i] = s.chain[--s.n]; return; }
}

void pdl_break_all() { g_pdl.n = 0; }

}  // namespace td3d
