// Gated-Delta-Net (Qwen3.5 linear-attention layers) kernels.  See gdn.cu.
#pragma once

#include "common.cuh"
#include "gdn_args.h"

namespace cb {

int gdn_forward_launch(cudaStream_t st, const GdnArgs& a);
int gdn_forward_launch_count(const GdnArgs& a);
// chunkwise recurrence (gdn_chunk.cu): a.glog and a.chunk_ws set, S >= GDN_CHUNK; replaces gdn_recur_kernel inside gdn_forward_launch
bool gdn_chunk_supported(const GdnArgs& a);
int gdn_chunk_recur_launch(cudaStream_t st, const GdnArgs& a);

}  // namespace cb
