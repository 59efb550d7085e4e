// fetch.cu -- IvfRabitqIndex::fetch_embedding (reference src/ivf.rs:1247-1307): reconstruct a stored vector from its codes.
//   find     which stored position holds the id (the reference scans the lists in order; ids are unique)
//   rebuild  binary code (simd::unpack_single_vector, src/simd.rs:915-960) + ex code (simd::unpack_ex_code, :101-134) ->
//            code = ex + (bit << ex_bits); rotated[i] = centroid[i] + delta * code + vl; inverse rotation
//            (FhtKacRotator::inverse_rotate_into src/rotation.rs:410-481, MatrixRotator :183-199).
// Not a hot path (one vector per call); it lives on the device because the codes do.  Bit-exact: element-wise steps and
// FHT butterflies are independent within a stage, the matrix rows are summed sequentially as the reference does.
#include "rotate.cuh"
#include "scan_common.cuh"

namespace rbq {

__global__ void find_id_kernel(const unsigned long long* __restrict__ ids, size_t n, unsigned long long id, unsigned long long* pos) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        if (ids[i] == id) atomicMin(pos, (unsigned long long)i);
}

template <int EXK>
__global__ void __launch_bounds__(256) fetch_embedding_kernel(DevIndex ix, uint32_t cid, uint32_t local, float delta, float vl,
                                                              float* __restrict__ out) {
    extern __shared__ float fe_smem[];
    float* buf = fe_smem;
    const int tid = threadIdx.x, nt = blockDim.x, D = ix.D;
    const uint8_t* blk = ix.blocks + ((size_t)ix.blk_off[cid] + local / kBatch) * ix.block_stride;
    const int v = (int)(local % kBatch), v16 = v & 15, hi = v >> 4;
    const int p = ((v16 & 7) << 1) | (v16 >> 3);  // KPERM0[p] = v16 (src/simd.rs:774)
    const uint8_t* ex = ix.ex + ((size_t)ix.vec_off[cid] + local) * ix.ex_stride;
    const float* cent = ix.centroids + (size_t)cid * D;
    for (int c = tid; c < D / 16; c += nt) {  // one 16-dim chunk per thread
        uint32_t X[4] = {0u, 0u, 0u, 0u};
        if (EXK != 0) decode_chunk<EXK, false>(ex, c, ix.ex_bits, X[0], X[1], X[2], X[3]);
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            const int d = 16 * c + r, cb = d >> 2;  // codebook cb holds dims 4cb..4cb+3, MSB first (src/simd.rs:141-150,771)
            const uint32_t byte = blk[16 * cb + p];
            const uint32_t nib = hi ? (byte >> 4) : (byte & 15u);
            const uint32_t bit = (nib >> (3 - (d & 3))) & 1u;
            const uint32_t code = ((X[r >> 2] >> (8 * (r & 3))) & 0xffu) + (bit << ix.ex_bits);
            const float m = delta * (float)code;
            const float a = cent[d] + m;
            buf[d] = a + vl;
        }
    }
    __syncthreads();
    if (ix.rot_type == RBQ_ROTATOR_FHT_KAC) {
        const int fo = D / 8;
        const bool pow2 = (ix.trunc == D);
        const int start = D - ix.trunc;
        const float inv_fac = 1.0f / ix.fac, inv_n = 1.0f / (float)ix.trunc;
        if (!pow2) {
            for (int i = tid; i < D; i += nt) buf[i] = buf[i] * 4.0f;
            __syncthreads();
        }
        for (int round = 3; round >= 0; --round) {
            if (!pow2) {
                const int half = D / 2;
                for (int i = tid; i < D; i += nt) buf[i] = buf[i] * 0.5f;
                __syncthreads();
                for (int i = tid; i < half; i += nt) {
                    const float x = buf[i], y = buf[i + half];
                    buf[i] = x + y;
                    buf[i + half] = x - y;
                }
                __syncthreads();
            }
            float* win = (pow2 || (round & 1) == 0) ? buf : buf + start;
            for (int i = tid; i < ix.trunc; i += nt) win[i] = win[i] * inv_fac;
            __syncthreads();
            fht_inplace(win, ix.trunc, tid, nt);
            for (int i = tid; i < ix.trunc; i += nt) win[i] = win[i] * inv_n;
            __syncthreads();
            const uint8_t* fl = ix.flip + round * fo;
            for (int i = tid; i < D; i += nt)
                if ((fl[i >> 3] >> (i & 7)) & 1) buf[i] = -buf[i];
            __syncthreads();
        }
        for (int i = tid; i < ix.dim; i += nt) out[i] = buf[i];
    } else {
        for (int col = tid; col < ix.dim; col += nt) {  // out = M^T rotated, rows summed in order
            float acc = 0.0f;
            const float* mt = ix.matrix_t + (size_t)col * D;  // matrix_t[col*D + row] = M[row][col]
            for (int row = 0; row < D; ++row) {
                const float pr = mt[row] * buf[row];
                acc = acc + pr;
            }
            out[col] = acc;
        }
    }
}

int launch_find_id(const DevIndex& ix, size_t nvec, uint64_t id, unsigned long long* d_pos, cudaStream_t st) {
    RBQ_CUDA(cudaMemsetAsync(d_pos, 0xff, 8, st));
    if (nvec == 0) return RBQ_OK;
    const unsigned grid = (unsigned)std::min<size_t>((nvec + 255) / 256, 1184);
    find_id_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const unsigned long long*>(ix.ids), nvec, (unsigned long long)id, d_pos);
    RBQ_CUDA(cudaGetLastError());
    return RBQ_OK;
}

int launch_fetch_embedding(const DevIndex& ix, uint32_t cid, uint32_t local, float delta, float vl, float* d_out, cudaStream_t st) {
    const size_t smem = (size_t)ix.D * 4;
    if (ix.ex_bits == 0) fetch_embedding_kernel<0><<<1, 256, smem, st>>>(ix, cid, local, delta, vl, d_out);
    else if (ix.ex_bits == 2) fetch_embedding_kernel<2><<<1, 256, smem, st>>>(ix, cid, local, delta, vl, d_out);
    else if (ix.ex_bits == 6) fetch_embedding_kernel<6><<<1, 256, smem, st>>>(ix, cid, local, delta, vl, d_out);
    else fetch_embedding_kernel<1><<<1, 256, smem, st>>>(ix, cid, local, delta, vl, d_out);
    RBQ_CUDA(cudaGetLastError());
    return RBQ_OK;
}

}  // namespace rbq
