// tcgen05 attention cores (attn_tc.cu): host-side problem description.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dclip {

// Both directions of the bidirectional cross-attention in one launch.  All matrices are 16-bit (bf16 / fp16) row-major:
//   Qt [B*T, d], KVi [B*R, 2d]  (direction 0: text queries over image keys / values)
//   Qi [B*R, d], KVt [B*T, 2d]  (direction 1)
// forward:  Ot [B*T, d], Oi [B*R, d]
// backward: dOt, dOi in; dQt, dKVi, dQi, dKVt out; optional per-batch column sums (in-projection bias gradient partials)
//           dbq_* [B, d], dbkv_* [B, 2d] in fp32.
struct AttnTcProblem {
    int B, T, R, d, heads, dtype16;
    const void* Qt; const void* KVi; const void* Qi; const void* KVt;
    void* Ot; void* Oi;
    const void* dOt; const void* dOi;
    void* dQt; void* dKVi; void* dQi; void* dKVt;
    float* dbq_t2i; float* dbkv_t2i; float* dbq_i2t; float* dbkv_i2t;
};

bool attn_tc_shape_ok(int64_t T, int64_t R, int64_t d, int heads);
int launch_attn_tc(const AttnTcProblem& p, bool backward, cudaStream_t st);

}  // namespace dclip
