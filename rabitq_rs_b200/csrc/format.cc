// format.cc -- "RBQ1" v3 index stream: parser, writer, CRC-32 and shard assignment (host code).
//
// Follows the reference's persistence format byte for byte (reference src/ivf.rs:1317-1474 save,
// :1484-1702 load): little-endian; magic + version outside the CRC; CRC-32/IEEE over everything
// between the version word and the trailer; same validation order and error strings.
#include <algorithm>
#include <cstring>
#include <mutex>

#include "rbq_internal.h"

namespace rbq {

static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }
int fail(int code, const std::string& msg) {
    static const char* prefix[] = {"", "dimension mismatch: ", "invalid configuration: ", "",
                                   "i/o error while reading or writing an index: ",
                                   "invalid persisted index: ", ""};
    g_last_error = std::string(code >= 0 && code <= 6 ? prefix[code] : "") + msg;
    return code;
}
const char* last_error_cstr() { return g_last_error.c_str(); }

// ---- CRC-32/IEEE, slice-by-8 ---------------------------------------------------------------
static uint32_t g_crc[8][256];
static std::once_flag g_crc_once;  // loads run outside any lock (one thread per GPU shard)
static void crc_init() {
    for (uint32_t i = 0; i < 256; ++i) {
        uint32_t c = i;
        for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
        g_crc[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; ++i)
        for (int t = 1; t < 8; ++t) g_crc[t][i] = (g_crc[t - 1][i] >> 8) ^ g_crc[0][g_crc[t - 1][i] & 0xff];
}
uint32_t crc32_ieee(uint32_t crc, const uint8_t* p, size_t n) {
    std::call_once(g_crc_once, crc_init);
    crc = ~crc;
    while (n >= 8) {
        uint32_t a, b;
        std::memcpy(&a, p, 4);
        std::memcpy(&b, p + 4, 4);
        a ^= crc;
        crc = g_crc[7][a & 0xff] ^ g_crc[6][(a >> 8) & 0xff] ^ g_crc[5][(a >> 16) & 0xff] ^ g_crc[4][a >> 24] ^
              g_crc[3][b & 0xff] ^ g_crc[2][(b >> 8) & 0xff] ^ g_crc[1][(b >> 16) & 0xff] ^ g_crc[0][b >> 24];
        p += 8;
        n -= 8;
    }
    while (n--) crc = g_crc[0][(crc ^ *p++) & 0xff] ^ (crc >> 8);
    return ~crc;
}

// Size-balanced list -> shard assignment: largest list first onto the least loaded shard
// (ties: lower shard id).  Deterministic, so every rank derives the same map from the same file.
void assign_shards(const std::vector<uint64_t>& list_bytes, int shard_count, std::vector<int>& owner) {
    size_t nl = list_bytes.size();
    owner.assign(nl, 0);
    if (shard_count <= 1) return;
    std::vector<uint32_t> order(nl);
    for (size_t i = 0; i < nl; ++i) order[i] = (uint32_t)i;
    std::stable_sort(order.begin(), order.end(),
                     [&](uint32_t a, uint32_t b) { return list_bytes[a] > list_bytes[b]; });
    std::vector<uint64_t> load(shard_count, 0);
    for (uint32_t c : order) {
        int best = 0;
        for (int s = 1; s < shard_count; ++s)
            if (load[s] < load[best]) best = s;
        owner[c] = best;
        load[best] += list_bytes[c] + 1;
    }
}

namespace {
struct Reader {
    const uint8_t* p;
    size_t n, off = 0;
    bool eof = false;
    bool get(void* d, size_t k) {
        if (k > n - off) {
            eof = true;
            off = n;
            if (d) std::memset(d, 0, k);
            return false;
        }
        if (d) std::memcpy(d, p + off, k);
        off += k;
        return true;
    }
    bool skip(size_t k) { return get(nullptr, k); }
    template <class T>
    T rd() {
        T v{};
        get(&v, sizeof(T));
        return v;
    }
};
}  // namespace

// From list_n_all + (shard_rank, shard_count): the owner of every list (size-balanced over the bytes a list occupies on the
// device) and this shard's geometry (list_n, blk_off, vec_off).  Shared by the loader and the streaming builder, so that a
// shard built in place equals the same shard loaded from the complete file.
int shard_layout(HostIndex& ix) {
    const uint64_t ncl = ix.nlist;
    const size_t stride = ix.block_stride(), exs = ix.ex_stride();
    std::vector<uint64_t> bytes(ncl);
    for (uint64_t c = 0; c < ncl; ++c) {
        uint64_t nv = ix.list_n_all[c];
        bytes[c] = (nv + kBatch - 1) / kBatch * stride + nv * (exs + 16);
    }
    std::vector<int> owner;
    assign_shards(bytes, ix.shard_count, owner);
    ix.list_owner.resize(ncl);
    for (uint64_t c = 0; c < ncl; ++c) ix.list_owner[c] = (uint8_t)owner[c];
    ix.list_n.assign(ncl, 0);
    ix.blk_off.assign(ncl + 1, 0);
    ix.vec_off.assign(ncl + 1, 0);
    uint64_t own_vec = 0, own_blk = 0;
    for (uint64_t c = 0; c < ncl; ++c) {
        ix.blk_off[c] = (uint32_t)own_blk;
        ix.vec_off[c] = own_vec;
        if (owner[c] == ix.shard_rank) {
            ix.list_n[c] = ix.list_n_all[c];
            own_vec += ix.list_n_all[c];
            own_blk += (ix.list_n_all[c] + kBatch - 1) / kBatch;
        }
    }
    ix.blk_off[ncl] = (uint32_t)own_blk;
    ix.vec_off[ncl] = own_vec;
    if (own_blk > 0xFFFFFFFFull) return fail(RBQ_INVALID_CONFIG, "shard holds more than 2^32 blocks");
    return RBQ_OK;
}

// Two passes over the stream: (1) walk it to learn every list's size (needed for the shard map)
// and validate structure; (2) copy the lists this shard owns.  The CRC covers the whole stream.
int parse_rbq1(const uint8_t* p, size_t n, int shard_rank, int shard_count, HostIndex& ix) {
    const char* kEof = "failed to fill whole buffer";
    if (shard_count < 1 || shard_rank < 0 || shard_rank >= shard_count)
        return fail(RBQ_INVALID_CONFIG, "shard_rank/shard_count out of range");
    if (shard_count > kMaxShards)  // list_owner is a byte per list and the device merge walks <= kMaxShards sorted lists
        return fail(RBQ_INVALID_CONFIG, "shard_count exceeds the supported maximum (32)");
    Reader r{p, n};
    char magic[4];
    if (!r.get(magic, 4)) return fail(RBQ_IO, kEof);
    if (std::memcmp(magic, "RBQ1", 4) != 0) return fail(RBQ_INVALID_PERSISTENCE, "unrecognized file header");
    uint32_t version = r.rd<uint32_t>();
    if (r.eof) return fail(RBQ_IO, kEof);
    if (version != 3)
        return fail(RBQ_INVALID_PERSISTENCE,
                    "unsupported index format version (expected V3 with unified memory layout)");
    uint32_t dim = r.rd<uint32_t>();
    if (r.eof) return fail(RBQ_IO, kEof);
    if (dim == 0) return fail(RBQ_INVALID_PERSISTENCE, "dimension must be positive");
    uint32_t D = r.rd<uint32_t>();
    if (r.eof) return fail(RBQ_IO, kEof);
    if (D < dim) return fail(RBQ_INVALID_PERSISTENCE, "padded_dim must be >= dim");
    uint8_t metric = r.rd<uint8_t>();
    if (r.eof) return fail(RBQ_IO, kEof);
    if (metric > 1) return fail(RBQ_INVALID_PERSISTENCE, "unknown metric tag");
    uint8_t rot = r.rd<uint8_t>();
    if (r.eof) return fail(RBQ_IO, kEof);
    if (rot > 1) return fail(RBQ_INVALID_PERSISTENCE, "unknown rotator type tag");
    uint8_t exb = r.rd<uint8_t>();
    if (r.eof) return fail(RBQ_IO, kEof);
    if (exb > 16) return fail(RBQ_INVALID_PERSISTENCE, "ex_bits out of range");
    uint8_t tb = r.rd<uint8_t>();
    if (r.eof) return fail(RBQ_IO, kEof);
    if (tb == 0 || tb > 16) return fail(RBQ_INVALID_PERSISTENCE, "total_bits out of range");
    if ((int)tb - 1 != (int)exb) return fail(RBQ_INVALID_PERSISTENCE, "total_bits does not match ex_bits");
    uint64_t nvec = r.rd<uint64_t>(), ncl = r.rd<uint64_t>(), rlen = r.rd<uint64_t>();
    if (r.eof) return fail(RBQ_IO, kEof);
    if (rlen > n - r.off) return fail(RBQ_IO, kEof);
    if (rot == 1 && rlen != 4 * (uint64_t)D / 8)
        return fail(RBQ_INVALID_PERSISTENCE, "FHT rotator flip bits length mismatch");
    if (rot == 0 && rlen != (uint64_t)D * D * 4)
        return fail(RBQ_INVALID_PERSISTENCE, "rotator matrix length mismatch");
    ix = HostIndex();
    ix.dim = dim;
    ix.D = D;
    ix.metric = metric;
    ix.rot_type = rot;
    ix.ex_bits = exb;
    ix.rot_bytes.assign(p + r.off, p + r.off + rlen);
    r.off += rlen;
    ix.shard_rank = shard_rank;
    ix.shard_count = shard_count;
    if (ncl > (n - r.off) / ((uint64_t)D * 4 + 16) + 1) return fail(RBQ_IO, kEof);
    ix.nlist = ncl;
    const size_t stride = ix.block_stride(), exs = ix.ex_stride();
    const size_t lists_begin = r.off;

    // pass 1
    ix.list_n_all.resize(ncl);
    std::vector<size_t> list_pos(ncl);
    uint64_t total = 0;
    for (uint64_t c = 0; c < ncl; ++c) {
        list_pos[c] = r.off;
        if (!r.skip((size_t)D * 4)) return fail(RBQ_IO, kEof);
        uint64_t nv = r.rd<uint64_t>();
        if (r.eof) return fail(RBQ_IO, kEof);
        if (nv > 1000000)
            return fail(RBQ_INVALID_PERSISTENCE, "cluster size exceeds reasonable limits - possible corruption");
        if (!r.skip(nv * 8)) return fail(RBQ_IO, kEof);
        uint64_t bl = r.rd<uint64_t>();
        if (r.eof) return fail(RBQ_IO, kEof);
        if (bl != stride * ((nv + kBatch - 1) / kBatch))
            return fail(RBQ_INVALID_PERSISTENCE,
                        "batch_data length mismatch - possible corruption or version incompatibility");
        if (!r.skip(bl)) return fail(RBQ_IO, kEof);
        for (uint64_t v = 0; v < nv; ++v) {
            uint64_t el = r.rd<uint64_t>();
            if (r.eof) return fail(RBQ_IO, kEof);
            if (el != exs)
                return fail(RBQ_INVALID_PERSISTENCE,
                            "ex_code_packed length mismatch - possible corruption or version incompatibility");
            if (!r.skip(exs)) return fail(RBQ_IO, kEof);
        }
        if (!r.skip(nv * 16)) return fail(RBQ_IO, kEof);
        ix.list_n_all[c] = (uint32_t)nv;
        total += nv;
    }
    if (total != nvec) return fail(RBQ_INVALID_PERSISTENCE, "vector count metadata mismatch");
    const size_t crc_end = r.off;
    uint32_t stored = r.rd<uint32_t>();
    if (r.eof) return fail(RBQ_IO, kEof);
    if (crc32_ieee(0, p + 8, crc_end - 8) != stored) return fail(RBQ_INVALID_PERSISTENCE, "checksum mismatch");
    ix.nvec_total = nvec;
    (void)lists_begin;

    int rc_layout = shard_layout(ix);
    if (rc_layout) return rc_layout;
    std::vector<int> owner(ix.list_owner.begin(), ix.list_owner.end());
    const uint64_t own_vec = ix.vec_off[ncl], own_blk = ix.blk_off[ncl];
    ix.centroids.resize(ncl * (size_t)D);
    ix.blocks.resize(own_blk * stride);
    ix.ids.resize(own_vec);
    ix.ex.resize(own_vec * exs);
    ix.f_add_ex.resize(own_vec);
    ix.f_rescale_ex.resize(own_vec);
    ix.delta.resize(own_vec);
    ix.vl.resize(own_vec);
    for (uint64_t c = 0; c < ncl; ++c) {
        const uint8_t* q = p + list_pos[c];
        std::memcpy(&ix.centroids[c * (size_t)D], q, (size_t)D * 4);
        if (owner[c] != shard_rank) continue;
        q += (size_t)D * 4 + 8;
        size_t nv = ix.list_n_all[c], vo = ix.vec_off[c];
        std::memcpy(ix.ids.data() + vo, q, nv * 8);
        q += nv * 8 + 8;
        size_t bl = (nv + kBatch - 1) / kBatch * stride;
        std::memcpy(ix.blocks.data() + (size_t)ix.blk_off[c] * stride, q, bl);
        q += bl;
        for (size_t v = 0; v < nv; ++v) {
            q += 8;
            if (exs) std::memcpy(ix.ex.data() + (vo + v) * exs, q, exs);
            q += exs;
        }
        std::memcpy(ix.f_add_ex.data() + vo, q, nv * 4);
        q += nv * 4;
        std::memcpy(ix.f_rescale_ex.data() + vo, q, nv * 4);
        q += nv * 4;
        std::memcpy(ix.delta.data() + vo, q, nv * 4);
        q += nv * 4;
        std::memcpy(ix.vl.data() + vo, q, nv * 4);
    }
    return RBQ_OK;
}

void write_rbq1(const HostIndex& ix, std::vector<uint8_t>& out) {
    out.clear();
    const size_t stride = ix.block_stride(), exs = ix.ex_stride();
    size_t total = 8 + 12 + 24 + ix.rot_bytes.size() + 4;
    for (size_t c = 0; c < ix.nlist; ++c) {
        size_t nv = ix.list_n[c];
        total += (size_t)ix.D * 4 + 16 + nv * 8 + (nv + kBatch - 1) / kBatch * stride + nv * (8 + exs) + nv * 16;
    }
    out.reserve(total);
    auto put = [&](const void* q, size_t k) {
        const uint8_t* b = (const uint8_t*)q;
        out.insert(out.end(), b, b + k);
    };
    auto u64 = [&](uint64_t v) { put(&v, 8); };
    put("RBQ1", 4);
    uint32_t v3 = 3;
    put(&v3, 4);
    put(&ix.dim, 4);
    put(&ix.D, 4);
    uint8_t tags[4] = {(uint8_t)ix.metric, (uint8_t)ix.rot_type, (uint8_t)ix.ex_bits, (uint8_t)(ix.ex_bits + 1)};
    put(tags, 4);
    u64(ix.nvec_total);
    u64(ix.nlist);
    u64(ix.rot_bytes.size());
    put(ix.rot_bytes.data(), ix.rot_bytes.size());
    for (size_t c = 0; c < ix.nlist; ++c) {
        size_t nv = ix.list_n[c], vo = ix.vec_off[c];
        put(&ix.centroids[c * (size_t)ix.D], (size_t)ix.D * 4);
        u64(nv);
        put(ix.ids.data() + vo, nv * 8);
        size_t bl = (nv + kBatch - 1) / kBatch * stride;
        u64(bl);
        put(ix.blocks.data() + (size_t)ix.blk_off[c] * stride, bl);
        for (size_t v = 0; v < nv; ++v) {
            u64(exs);
            if (exs) put(ix.ex.data() + (vo + v) * exs, exs);
        }
        put(ix.f_add_ex.data() + vo, nv * 4);
        put(ix.f_rescale_ex.data() + vo, nv * 4);
        put(ix.delta.data() + vo, nv * 4);
        put(ix.vl.data() + vo, nv * 4);
    }
    uint32_t crc = crc32_ieee(0, out.data() + 8, out.size() - 8);
    put(&crc, 4);
}

}  // namespace rbq

extern "C" const char* rbq_last_error(void) { return rbq::last_error_cstr(); }

extern "C" int rbq_shard_assignment(const uint8_t* bytes, size_t len, int shard_count, int32_t* owner, uint32_t* list_sizes,
                                    size_t cap_lists, size_t* nlist_out) {
    if (!bytes || !nlist_out) return rbq::fail(RBQ_INVALID_CONFIG, "null argument");
    rbq::HostIndex hi;
    // parse as shard 0 of shard_count: validates the stream and yields every list's size
    int rc = rbq::parse_rbq1(bytes, len, 0, shard_count, hi);
    if (rc) return rc;
    *nlist_out = hi.nlist;
    if (cap_lists < hi.nlist) return owner ? rbq::fail(RBQ_INVALID_CONFIG, "output buffer too small") : (int)RBQ_OK;
    std::vector<uint64_t> bytes_per_list(hi.nlist);
    const size_t stride = hi.block_stride(), exs = hi.ex_stride();
    for (size_t c = 0; c < hi.nlist; ++c) {
        const uint64_t nv = hi.list_n_all[c];
        bytes_per_list[c] = (nv + rbq::kBatch - 1) / rbq::kBatch * stride + nv * (exs + 16);
        if (list_sizes) list_sizes[c] = (uint32_t)nv;
    }
    std::vector<int> own;
    rbq::assign_shards(bytes_per_list, shard_count, own);
    if (owner)
        for (size_t c = 0; c < hi.nlist; ++c) owner[c] = own[c];
    return RBQ_OK;
}
