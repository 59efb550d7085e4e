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
#include <cstdlib>

#include "rbq_internal.h"
#include <cuda_bf16.h>

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

// ---- warp-per-query variant (FhtKac rotator, trunc >= 64) -----------------------------------------------
// One warp owns one query: no block barriers, the FHT window lives in registers (element r = e*32 + lane of
// the window): strides 1..16 are lane exchanges (shfl.xor), strides >= 32 pair registers.  Stages run in the
// reference's order (h = 1, 2, 4, ...) and every butterfly is the reference's (x + y, x - y), so the result is
// bit-identical to fht() (reference src/rotation.rs:292-313).  x - y is evaluated as x + (-y) (same IEEE value).
constexpr int kPrepWarps = 4;

// FASTQ: the LUT quantiser's `round((t - vl) / delta)` (an IEEE division + roundf per entry, ~20 instructions) is evaluated as
// t' = (t - vl) * fl(1 / delta), whose distance from the reference's quotient is below 4.6e-5 for quotients in [0, 256)
// (one rounding of the reciprocal, one of the product, half an ulp of the true quotient), and rounded with the 2^23 trick; an
// entry whose t' lies within 1e-4 of a rounding boundary (or is not finite) takes the reference's expression instead, so the
// byte is the reference's in every case.
template <int E, bool FASTQ>
__global__ void __launch_bounds__(kPrepWarps * 32) query_prep_fht_kernel(DevIndex ix, const float* __restrict__ queries, uint32_t nq,
                                                                         float* __restrict__ rot_out, uint8_t* __restrict__ lut_out,
                                                                         QueryScalars* __restrict__ qs_out,
                                                                         __nv_bfloat16* __restrict__ split_out, float* __restrict__ n2_out) {
    extern __shared__ __align__(16) float smem[];
    const int D = ix.D, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t q = blockIdx.x * kPrepWarps + warp;
    if (q >= nq) return;  // warp-uniform; the kernel has no block-wide barrier
    float* buf = smem + (size_t)warp * 2 * D;  // D rotated values
    float* sq = buf + D;                       // D squares (second sequential fold)
    const float* qin = queries + (size_t)q * ix.dim;
    for (int i = lane; i < D; i += 32) buf[i] = i < ix.dim ? __ldg(qin + i) : 0.0f;
    __syncwarp();

    const bool pow2 = (ix.trunc == D);
    const int start = D - ix.trunc, half = D / 2;
    for (int round = 0; round < 4; ++round) {
        const uint32_t* fw = reinterpret_cast<const uint32_t*>(ix.flip + round * (D / 8));  // bit i%32 of word i/32 (LSB first)
        const int base = (pow2 || (round & 1) == 0) ? 0 : start;                            // multiple of 32
        float v[E];
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int idx = base + e * 32 + lane;
            const uint32_t sgn = ((__ldg(fw + (idx >> 5)) >> lane) & 1u) << 31;
            v[e] = __uint_as_float(__float_as_uint(buf[idx]) ^ sgn);
        }
        if (!pow2) {  // the sign flip covers the whole padded vector, not only the FHT window
            for (int i = lane; i < D; i += 32)
                if (i < base || i >= base + ix.trunc) {
                    const uint32_t sgn = ((__ldg(fw + (i >> 5)) >> lane) & 1u) << 31;
                    buf[i] = __uint_as_float(__float_as_uint(buf[i]) ^ sgn);
                }
        }
#pragma unroll
        for (int m = 1; m < 32; m <<= 1) {
            const uint32_t neg = (lane & m) ? 0x80000000u : 0u;
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const float other = __shfl_xor_sync(0xffffffffu, v[e], m);
                v[e] = other + __uint_as_float(__float_as_uint(v[e]) ^ neg);
            }
        }
#pragma unroll
        for (int s = 1; s < E; s <<= 1) {
#pragma unroll
            for (int e = 0; e < E; ++e)
                if ((e & s) == 0) {
                    const float x = v[e], y = v[e + s];
                    v[e] = x + y;
                    v[e + s] = x - y;
                }
        }
#pragma unroll
        for (int e = 0; e < E; ++e) buf[base + e * 32 + lane] = v[e] * ix.fac;
        __syncwarp();
        if (!pow2) {  // kacs_walk over the whole padded vector
            for (int i = lane; i < half; i += 32) {
                const float x = buf[i], y = buf[i + half];
                buf[i] = x + y;
                buf[i + half] = x - y;
            }
            __syncwarp();
        }
    }
    for (int i = lane; i < D; i += 32) {
        float x = buf[i];
        if (!pow2) {
            x = x * 0.25f;
            buf[i] = x;
        }
        sq[i] = x * x;
        rot_out[(size_t)q * D + i] = x;
        if (split_out != nullptr) {  // operand of the tensor-core coarse stage (coarse_tc.cu): [hi | hi | lo] bf16 split
            const __nv_bfloat16 hi = __float2bfloat16_rn(x);
            const __nv_bfloat16 lo = __float2bfloat16_rn(x - __bfloat162float(hi));
            __nv_bfloat16* o = split_out + (size_t)q * 3 * D;
            o[i] = hi;
            o[D + i] = hi;
            o[2 * D + i] = lo;
        }
    }
    __syncwarp();

    // K3: the reference folds sum(q) and sum(q*q) sequentially (ivf.rs:863-864): lane 0 / lane 1 walk the arrays in order
    float fold = 0.0f;
    {
        const float4* src = reinterpret_cast<const float4*>(lane == 1 ? sq : buf);
        if (lane < 2) {
#pragma unroll 4
            for (int i = 0; i < D / 4; ++i) {
                const float4 x = src[i];
                fold = fold + x.x;
                fold = fold + x.y;
                fold = fold + x.z;
                fold = fold + x.w;
            }
        }
    }
    const float s_sum = __shfl_sync(0xffffffffu, fold, 0), s_sumsq = __shfl_sync(0xffffffffu, fold, 1);

    // K2: float LUT of codebook cb (dims 4cb..4cb+3): entry j = lut[j - lowbit(j)] + q[4cb + KPOS[j]]
    const int ncb = D / 4;
    auto table = [&](int cb, float (&t)[16]) {
        const float4 qv = *reinterpret_cast<const float4*>(buf + 4 * cb);
        t[0] = 0.0f;
        t[1] = t[0] + qv.w;
        t[2] = t[0] + qv.z;
        t[3] = t[2] + qv.w;
        t[4] = t[0] + qv.y;
        t[5] = t[4] + qv.w;
        t[6] = t[4] + qv.z;
        t[7] = t[6] + qv.w;
        t[8] = t[0] + qv.x;
        t[9] = t[8] + qv.w;
        t[10] = t[8] + qv.z;
        t[11] = t[10] + qv.w;
        t[12] = t[8] + qv.y;
        t[13] = t[12] + qv.w;
        t[14] = t[12] + qv.z;
        t[15] = t[14] + qv.w;
    };
    float lmin = FLT_MAX, lmax = -FLT_MAX;
    for (int cb = lane; cb < ncb; cb += 32) {
        float t[16];
        table(cb, t);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            lmin = fminf(lmin, t[j]);
            lmax = fmaxf(lmax, t[j]);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lmin = fminf(lmin, __shfl_xor_sync(0xffffffffu, lmin, o));
        lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
    }
    const float vl = lmin, vr = lmax;
    const float delta = (vr - vl) / 255.0f;
    uint4* lut128 = reinterpret_cast<uint4*>(lut_out + (size_t)q * D * 4);
    const float rinv = 1.0f / delta;
    for (int cb = lane; cb < ncb; cb += 32) {
        uint32_t w[4] = {0u, 0u, 0u, 0u};
        if (delta > 0.0f) {
            float t[16];
            table(cb, t);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                uint32_t byte;
                const float a = t[j] - vl;
                const float y = a * rinv + 8388608.0f;     // separate mul and add (-fmad=false); integer part in the low mantissa bits
                const float d = a * rinv - (y - 8388608.0f);
                if (FASTQ && fabsf(d) <= 0.4999f && y < 8388864.0f) {  // false for NaN / inf as well
                    byte = __float_as_uint(y) & 0x1ffu;
                    byte = byte > 255u ? 255u : byte;
                } else {
                    float x = roundf(a / delta);
                    x = fminf(fmaxf(x, 0.0f), 255.0f);  // clamp(0,255); NaN -> 0 like `as u8`
                    byte = (uint32_t)(x == x ? (int)x : 0);
                }
                w[j >> 2] |= byte << (8 * (j & 3));
            }
        }
        lut128[cb] = make_uint4(w[0], w[1], w[2], w[3]);
    }
    if (lane == 0) {
        QueryScalars s;
        const float bscale = (float)(1 << ix.ex_bits);
        const float cbv = -(bscale - 0.5f);
        s.delta = delta;
        s.sum_vl = vl * (float)ncb;
        s.k1x = -0.5f * s_sum;
        s.kbx = cbv * s_sum;
        s.qnorm = sqrtf(s_sumsq);
        s.sum_q = s_sum;
        s.bscale = bscale;
        s.pad = 0.0f;
        qs_out[q] = s;
        if (n2_out != nullptr) n2_out[q] = s_sumsq;  // |q|^2 for the approximate scores (any summation order will do)
    }
}

template <int E, bool FASTQ>
static int launch_prep_fht_q(const DevIndex& ix, const float* d_queries, size_t nq, float* d_rot, uint8_t* d_lut, QueryScalars* d_qs,
                           cudaStream_t st, void* d_split, float* d_n2) {
    const size_t smem = (size_t)kPrepWarps * 2 * ix.D * sizeof(float);
    if (smem > 48 * 1024)
        RBQ_CUDA(cudaFuncSetAttribute(query_prep_fht_kernel<E, FASTQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    query_prep_fht_kernel<E, FASTQ><<<(unsigned)((nq + kPrepWarps - 1) / kPrepWarps), kPrepWarps * 32, smem, st>>>(ix, d_queries, (uint32_t)nq,
                                                                                                            d_rot, d_lut, d_qs,
                                                                                                            reinterpret_cast<__nv_bfloat16*>(d_split), d_n2);
    RBQ_CUDA(cudaGetLastError());
    return RBQ_OK;
}

constexpr int kPrepFastQuantDefault = 1;
template <int E>
static int launch_prep_fht(const DevIndex& ix, const float* d_queries, size_t nq, float* d_rot, uint8_t* d_lut, QueryScalars* d_qs,
                           cudaStream_t st, void* d_split, float* d_n2) {
    const char* e = getenv("RBQ_PREP_FASTQ");  // tuning knob, read per launch
    if (e ? atoi(e) != 0 : kPrepFastQuantDefault != 0) return launch_prep_fht_q<E, true>(ix, d_queries, nq, d_rot, d_lut, d_qs, st, d_split, d_n2);
    return launch_prep_fht_q<E, false>(ix, d_queries, nq, d_rot, d_lut, d_qs, st, d_split, d_n2);
}

int launch_query_prep(const DevIndex& ix, const float* d_queries, size_t nq, float* d_rot, uint8_t* d_lut,
                      QueryScalars* d_qs, cudaStream_t st, void* d_split, float* d_n2, bool* split_done) {
    if (split_done) *split_done = false;
    if (nq == 0) return RBQ_OK;
    if (ix.rot_type == RBQ_ROTATOR_FHT_KAC && ix.trunc >= 64 && ix.trunc <= 2048) {
        if (split_done) *split_done = d_split != nullptr && d_n2 != nullptr;
        if (d_split == nullptr || d_n2 == nullptr) d_split = nullptr, d_n2 = nullptr;
        switch (ix.trunc / 32) {
            case 2: return launch_prep_fht<2>(ix, d_queries, nq, d_rot, d_lut, d_qs, st, d_split, d_n2);
            case 4: return launch_prep_fht<4>(ix, d_queries, nq, d_rot, d_lut, d_qs, st, d_split, d_n2);
            case 8: return launch_prep_fht<8>(ix, d_queries, nq, d_rot, d_lut, d_qs, st, d_split, d_n2);
            case 16: return launch_prep_fht<16>(ix, d_queries, nq, d_rot, d_lut, d_qs, st, d_split, d_n2);
            case 32: return launch_prep_fht<32>(ix, d_queries, nq, d_rot, d_lut, d_qs, st, d_split, d_n2);
            default: return launch_prep_fht<64>(ix, d_queries, nq, d_rot, d_lut, d_qs, st, d_split, d_n2);
        }
    }
    size_t smem = (size_t)ix.D * 5 * sizeof(float);
    if (smem > 48 * 1024)
        RBQ_CUDA(cudaFuncSetAttribute(query_prep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int threads = ix.D >= 512 ? 256 : 128;
    query_prep_kernel<<<(unsigned)nq, threads, smem, st>>>(ix, d_queries, d_rot, d_lut, d_qs);
    RBQ_CUDA(cudaGetLastError());
    return RBQ_OK;
}

}  // namespace rbq
