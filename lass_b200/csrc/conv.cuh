// Host-side description of one implicit-GEMM convolution launch (kernels K3 / K4, conv.cu).
#pragma once
#include "lass_internal.cuh"

namespace lass {

// The launch description IS the public C struct (include/lass_b200.h).
typedef lass_conv_segment ConvSegment;
typedef lass_conv_out ConvOut;
typedef lass_conv_desc ConvLaunch;

// Prepared launch: tensor maps encoded, tile configuration chosen.
struct ConvPrepared;
int conv_prepare(const ConvLaunch& l, ConvPrepared** out);
int conv_run(const ConvPrepared* p, cudaStream_t stream);
void conv_free(ConvPrepared* p);
double conv_flops(const ConvLaunch& l);
void conv_set_debug_flags(int flags);
void conv_set_profile_buffer(long long* buf);  // device buffer, >= grid * 16 int64; nullptr = off

}  // namespace lass
