// query_prep.cu -- per-query preparation on the device, one CTA per query:
//   K1  rotation          FhtKacRotator::rotate_into   (reference src/rotation.rs:350-401)
//                         MatrixRotator::rotate_into   (reference src/rotation.rs:158-173)
//   K3  query constants   QueryPrecomputed::new        (reference src/ivf.rs:862-878)
//   K2  FastScan u8 LUT   QueryLut::new + pack_lut_f32 (reference src/ivf.rs:798-845, src/simd.rs:818-840)
//
// Bit-exactness: every float operation is the reference's, in the reference's order.  The file is
// compiled with -fmad=false so no mul+add pair is contracted; the butterflies of one FHT stage are
// independent, so running them in parallel does not change any rounding.  Division is IEEE
// (-prec-div=true); roundf is round-half-away-from-zero like f32::round.
#include <cfloat>

#include "rbq_internal.h"
#include "rotate.cuh"

namespace rbq {

__global__ void __launch_bounds__(256) query_prep_kernel(DevIndex ix, const float* __restrict__ queries,
                                                         float* __restrict__ rot_out, uint8_t* __restrict__ lut_out,
                                                         QueryScalars* __restrict__ qs_out) {
    extern __shared__ float smem[];
    const int D = ix.D, tid = threadIdx.x, nt = blockDim.x;
    float* buf = smem;           // D
    float* lutf = smem + D;      // 4*D
    __shared__ float red_min[8], red_max[8];
    __shared__ float s_sum, s_sumsq;
    const size_t q = blockIdx.x;
    const float* qin = queries + q * ix.dim;

    rotate_block(ix, qin, buf, lutf, tid, nt);

    for (int i = tid; i < D; i += nt) rot_out[q * D + i] = buf[i];

    // K3: rotated_query.iter().sum() and sum of squares are sequential folds in the reference
    // (ivf.rs:863-864); one lane each reproduces the exact association order.
    if (tid == 0) {
        float s = 0.0f;
        for (int i = 0; i < D; ++i) s = s + buf[i];
        s_sum = s;
    } else if (tid == 32) {
        float s = 0.0f;
        for (int i = 0; i < D; ++i) {
            float t = buf[i] * buf[i];
            s = s + t;
        }
        s_sumsq = s;
    }

    // K2: float LUT, one thread per 4-dim codebook; entry j = lut[j - lowbit(j)] + q[4i + KPOS[j]]
    const int ncb = D / 4;
    float lmin = FLT_MAX, lmax = -FLT_MAX;
    for (int cb = tid; cb < ncb; cb += nt) {
        const float q0 = buf[4 * cb], q1 = buf[4 * cb + 1], q2 = buf[4 * cb + 2], q3 = buf[4 * cb + 3];
        float t[16];
        t[0] = 0.0f;
        t[1] = t[0] + q3;   // KPOS[1]=3
        t[2] = t[0] + q2;   // KPOS[2]=2
        t[3] = t[2] + q3;
        t[4] = t[0] + q1;   // KPOS[4]=1
        t[5] = t[4] + q3;
        t[6] = t[4] + q2;
        t[7] = t[6] + q3;
        t[8] = t[0] + q0;   // KPOS[8]=0
        t[9] = t[8] + q3;
        t[10] = t[8] + q2;
        t[11] = t[10] + q3;
        t[12] = t[8] + q1;
        t[13] = t[12] + q3;
        t[14] = t[12] + q2;
        t[15] = t[14] + q3;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            lutf[cb * 16 + j] = t[j];
            lmin = fminf(lmin, t[j]);
            lmax = fmaxf(lmax, t[j]);
        }
    }
    // NaN queries: fminf/fmaxf drop NaNs where the reference's total_cmp would pick one; either way
    // every distance of such a query is non-finite and the result is empty (documented).
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lmin = fminf(lmin, __shfl_xor_sync(0xffffffffu, lmin, o));
        lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
    }
    if ((tid & 31) == 0) {
        red_min[tid >> 5] = lmin;
        red_max[tid >> 5] = lmax;
    }
    __syncthreads();
    float vl = red_min[0], vr = red_max[0];
    for (int w = 1; w < (nt >> 5); ++w) {
        vl = fminf(vl, red_min[w]);
        vr = fmaxf(vr, red_max[w]);
    }
    const float delta = (vr - vl) / 255.0f;
    uint32_t* lut32 = reinterpret_cast<uint32_t*>(lut_out + q * (size_t)D * 4);
    for (int w = tid; w < D; w += nt) {  // 4 entries per thread-iteration -> one 32-bit store
        uint32_t packed = 0;
        if (delta > 0.0f) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float v = roundf((lutf[4 * w + k] - vl) / delta);
                v = fminf(fmaxf(v, 0.0f), 255.0f);  // clamp(0,255); NaN -> 0 like `as u8`
                packed |= (uint32_t)(v == v ? (int)v : 0) << (8 * k);
            }
        }
        lut32[w] = packed;
    }
    if (tid == 0) {
        QueryScalars s;
        const float bscale = (float)(1 << ix.ex_bits);
        const float cb = -(bscale - 0.5f);
        s.delta = delta;
        s.sum_vl = vl * (float)ncb;
        s.k1x = -0.5f * s_sum;
        s.kbx = cb * s_sum;
        s.qnorm = sqrtf(s_sumsq);
        s.sum_q = s_sum;
        s.bscale = bscale;
        s.pad = 0.0f;
        qs_out[q] = s;
    }
}

int launch_query_prep(const DevIndex& ix, const float* d_queries, size_t nq, float* d_rot, uint8_t* d_lut,
                      QueryScalars* d_qs, cudaStream_t st) {
    if (nq == 0) return RBQ_OK;
    size_t smem = (size_t)ix.D * 5 * sizeof(float);
    if (smem > 48 * 1024)
        RBQ_CUDA(cudaFuncSetAttribute(query_prep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int threads = ix.D >= 512 ? 256 : 128;
    query_prep_kernel<<<(unsigned)nq, threads, smem, st>>>(ix, d_queries, d_rot, d_lut, d_qs);
    RBQ_CUDA(cudaGetLastError());
    return RBQ_OK;
}

}  // namespace rbq
