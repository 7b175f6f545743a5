// Internal interface of the tensor-core exemplar-prior forward (prior_lse_tc.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace exvae {

struct PriorTcArgs {
  const float* zp;   // [2, Bpad, KP]
  const float* mp;   // [2, Cpad, KP]
  int Bpad, Cpad, KP, B, C;
  const int* mcnt;   // NULL => no mask
  const int* mlist;
  const int64_t* cidx;
  const int64_t* z_idx;
  float* part;       // [Bpad, nsplit, 4]
};

bool prior_tc_enabled();
// launches the kernel; *nsplit_out = number of partials written per row
int prior_fwd_tc_launch(const PriorTcArgs& a, int* nsplit_out, cudaStream_t st);

}  // namespace exvae
