// Stages not implemented on the device yet report BS_ERR_UNSUPPORTED (never a CPU fallback).
#include "bs_common.cuh"
bs_status bs_dc_impl(const bs_volume* v, float, const float**, size_t*) { return bs_fail(v->ctx, BS_ERR_UNSUPPORTED, "dual contouring: not implemented yet"); }
