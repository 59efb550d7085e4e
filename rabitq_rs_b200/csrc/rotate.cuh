// rotate.cuh -- block-cooperative rotation of one vector held in shared memory.
//   FhtKacRotator::rotate_into (reference src/rotation.rs:350-401; flip_sign :278, fht :292,
//   kacs_walk :315, rescale :327) and MatrixRotator::rotate_into (reference src/rotation.rs:158-173).
// Bit-exact: the butterflies of one FHT stage are independent, so evaluating them in parallel
// changes no rounding; stages run in the reference's order (h = 1, 2, 4, ...).  Needs -fmad=false.
#pragma once
#include "rbq_internal.h"

namespace rbq {

__device__ __forceinline__ void fht_inplace(float* w, int n, int tid, int nthreads) {
    for (int h = 1; h < n; h <<= 1) {
        for (int idx = tid; idx < n / 2; idx += nthreads) {
            const int i = ((idx / h) * 2 * h) + (idx % h);
            const float x = w[i], y = w[i + h];
            w[i] = x + y;
            w[i + h] = x - y;
        }
        __syncthreads();
    }
}

// in: `dim` floats in global memory; buf: D floats of shared memory (result); tmp: D floats of shared
// scratch (Matrix rotator only).  All threads of the block must call; ends with a __syncthreads().
__device__ __forceinline__ void rotate_block(const DevIndex& ix, const float* __restrict__ in, float* buf, float* tmp,
                                             int tid, int nt) {
    const int D = ix.D, dim = ix.dim;
    if (ix.rot_type == RBQ_ROTATOR_FHT_KAC) {
        for (int i = tid; i < D; i += nt) buf[i] = i < dim ? in[i] : 0.0f;
        __syncthreads();
        const int fo = D / 8;
        const bool pow2 = (ix.trunc == D);
        const int start = D - ix.trunc;
        for (int round = 0; round < 4; ++round) {
            const uint8_t* fl = ix.flip + round * fo;
            for (int i = tid; i < D; i += nt)  // bit i%8 of byte i/8, LSB first
                if ((fl[i >> 3] >> (i & 7)) & 1) buf[i] = -buf[i];
            __syncthreads();
            float* win = (pow2 || (round & 1) == 0) ? buf : buf + start;
            fht_inplace(win, ix.trunc, tid, nt);
            for (int i = tid; i < ix.trunc; i += nt) win[i] = win[i] * ix.fac;
            __syncthreads();
            if (!pow2) {  // kacs_walk over the whole padded vector
                const int half = D / 2;
                for (int i = tid; i < half; i += nt) {
                    const float x = buf[i], y = buf[i + half];
                    buf[i] = x + y;
                    buf[i + half] = x - y;
                }
                __syncthreads();
            }
        }
        if (!pow2) {
            for (int i = tid; i < D; i += nt) buf[i] = buf[i] * 0.25f;
            __syncthreads();
        }
    } else {
        // out[row] = sequential sum_k pad[k] * M[row][k]; matrix_t is k-major so that consecutive
        // threads (rows) read consecutive floats
        for (int i = tid; i < D; i += nt) tmp[i] = i < dim ? in[i] : 0.0f;
        __syncthreads();
        for (int row = tid; row < D; row += nt) {
            float acc = 0.0f;
            for (int k = 0; k < D; ++k) {
                const float p = tmp[k] * ix.matrix_t[(size_t)k * D + row];
                acc = acc + p;
            }
            buf[row] = acc;
        }
        __syncthreads();
    }
}

}  // namespace rbq
