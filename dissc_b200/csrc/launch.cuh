// Host-side launch helpers shared between translation units (defined in generator.cu).
#pragma once
#include <vector>

#include "common.cuh"
#include "conv1d.cuh"

namespace dissc {

// -> [co_tile][chunk][CI_CHUNK][k][CO_TILE], zero padded.  transposed: source is (Cin,Cout,k).
std::vector<float> pack_weights(const float* w, int Cin, int Cout, int k, int co_tile, int ci_chunk, bool transposed);
int conv_co_tile(int Cout);
int conv_ci_chunk(int co_tile);
bool conv_supported(int k, int dil);
// fp32 CUDA-core fused conv (conv1d.cuh); p.w must be packed with pack_weights(conv_co_tile(Cout), conv_ci_chunk(..)).
int launch_conv(const ConvParams& p, int k, int dil, int co_tile, bool emb, cudaStream_t st);


// Per-handle error flags for out-of-range embedding ids (common.cuh::checked_row): two ints ([0] unit ids, [1] speaker
// ids) in mapped pinned host memory.
struct ErrFlag {
  int* host = nullptr;  // read by the host after a synchronisation
  int* dev = nullptr;   // the same word as seen from the device
};
int err_flag_create(ErrFlag* f);
void err_flag_destroy(ErrFlag* f);
// DISSC_OK, or DISSC_EINDEX (flag cleared) with a message naming the table(s) and their sizes
int err_flag_take(ErrFlag* f, const char* what, int unit_rows, int spkr_rows);

}  // namespace dissc
