// Rule-book (kernel map / indice pair) construction.  SURVEY.md section 8 rows a4, a8, a12.
//
// Layout is output-stationary and tiled for the convolution kernels in conv.cu / conv_tc.cu:
//   * output rows are cut into tiles of TM consecutive rows,
//   * tile t owns the fixed slab entries[t*TM*K, (t+1)*TM*K) -- no global scan, no atomics, the
//     layout is a pure function of the inputs (deterministic); only the occupied prefix of a slab
//     is ever touched, so DRAM traffic is 4 bytes per pair (the COO form is 8),
//   * inside a tile the pairs are bucketed by kernel offset k (seg[t][k] .. seg[t][k+1]) and, inside
//     a bucket, ordered by output row.  Within one bucket every output row appears at most once,
//     which is what lets the convolution accumulate a bucket into its shared-memory tile without
//     atomics.
//   * entry = (out_row_in_tile << 25) | in_row.
// One block builds one tile: TM*K hash probes (16-byte slot loads, L2 resident table) staged in
// shared memory, then one warp per offset compacts its bucket with ballots.
//
// The probe loop is instruction bound (ncu: profiles/r01_conv_v2_sass_notes.md), so it is kept minimal:
// the packed 64-bit key is LINEAR in the coordinates, hence key(in) = key(row base) + delta[k] with a per-offset
// delta precomputed once per block; lanes run over offsets (no division), rows over warps.
#include "common.cuh"

#define RB_THREADS 256

extern "C" int64_t insmos_rulebook_entries_capacity(int64_t n_out, int32_t K, int32_t TM) {
    if (n_out <= 0 || K <= 0 || TM <= 0) return 0;
    return ceil_div64(n_out, TM) * (int64_t)TM * K;
}

// coordinates may move by at most this many voxels through a kernel offset; rows whose base coordinate is closer
// than this to the edge of the packable range take the checked slow path
#define RB_GUARD 512

__device__ __forceinline__ int64_t packed_delta(int d0, int d1, int d2, int d3) {
    return (int64_t)d0 + (int64_t)d1 * 65536ll + (int64_t)d2 * 4294967296ll + (int64_t)d3 * 281474976710656ll;
}

// ------------------------------------------------------------------------------------------------
// X-BLOCK TABLE.  The cube maps probe runs of 3 or 5 consecutive x for every (row, other-dims offset): with the voxel
// table every probe is an independent random 16-byte access (ncu: the map build is bound by the divergent probe
// loop, ~100 M probes per forward).  A second table keyed by (x_index >> 2, y, z, t) holds the rows of the 4 voxels of
// an x-block in one 32-byte slot, so a run costs 1-2 slot lookups instead of 3-5.  x_index = floor(x / xstep) with
// xstep = tensor stride of the set (coordinates of a strided level are multiples of it).
struct __align__(32) XBlockSlot { unsigned long long key; int32_t rows[4]; unsigned long long pad; };

__global__ void k_xblock_clear(XBlockSlot* table, int64_t cap) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < cap) {
        int4* q = reinterpret_cast<int4*>(table + i);
        q[0] = make_int4(-1, -1, -1, -1);                          // key = all ones, rows[0..1] = -1
        q[1] = make_int4(-1, -1, 0, 0);                            // rows[2..3] = -1
    }
}
__global__ void k_xblock_insert(const int32_t* __restrict__ coords, int64_t n, int ncol, int xstep,
                                XBlockSlot* table, uint64_t mask) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t* c = coords + i * ncol;
    const int xi = floor_div(c[1], xstep);
    const uint64_t key = pack_key(c[0], xi >> 2, ncol > 2 ? c[2] : 0, ncol > 3 ? c[3] : 0, ncol > 4 ? c[4] : 0);
    uint64_t slot = hash64(key) & mask;
    while (true) {
        unsigned long long* kp = &table[slot].key;
        const unsigned long long cur = *reinterpret_cast<volatile unsigned long long*>(kp);
        if (cur == key) break;
        if (cur == INSMOS_EMPTY_KEY) {
            const unsigned long long prev = atomicCAS(kp, INSMOS_EMPTY_KEY, (unsigned long long)key);
            if (prev == INSMOS_EMPTY_KEY || prev == key) break;
        }
        slot = (slot + 1) & mask;
    }
    table[slot].rows[xi & 3] = (int32_t)i;
}
extern "C" int64_t insmos_xblock_capacity(int64_t n) {
    int64_t cap = 1024;
    while (cap < 2 * n) cap <<= 1;                                 // slots hold up to 4 voxels: load <= 0.5 even with 1 voxel per block
    return cap;
}
extern "C" int insmos_xblock_build(const int32_t* coords, int64_t n, int32_t ncol, int32_t xstep,
                                   void* table, int64_t cap, void* stream) {
    if (!table || cap <= 0 || (cap & (cap - 1)) || (n > 0 && !coords) || n < 0 || xstep < 1 || ncol < 2 || ncol > 5)
        return INSMOS_ERR_INVALID_ARG;
    if (cap < 2 * n) return INSMOS_ERR_INVALID_ARG;
    k_xblock_clear<<<(unsigned)ceil_div64(cap, 256), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<XBlockSlot*>(table), cap);
    INSMOS_CHECK_LAUNCH("k_xblock_clear");
    if (n > 0) {
        k_xblock_insert<<<(unsigned)ceil_div64(n, 256), 256, 0, (cudaStream_t)stream>>>(
            coords, n, ncol, xstep, reinterpret_cast<XBlockSlot*>(table), (uint64_t)(cap - 1));
        INSMOS_CHECK_LAUNCH("k_xblock_insert");
    }
    return INSMOS_OK;
}
// rows of the x-block with packed key `key` (all -1 when the block is absent); two 16-byte loads per probe step
__device__ __forceinline__ int4 xblock_find(const XBlockSlot* __restrict__ table, uint64_t mask, uint64_t key) {
    uint64_t slot = hash64(key) & mask;
    while (true) {
        const int4 v0 = __ldg(reinterpret_cast<const int4*>(table + slot));
        const uint64_t k = ((uint64_t)(uint32_t)v0.y << 32) | (uint32_t)v0.x;
        if (k == key) {
            const int4 v1 = __ldg(reinterpret_cast<const int4*>(table + slot) + 1);
            return make_int4(v0.z, v0.w, v1.x, v1.y);
        }
        if (k == INSMOS_EMPTY_KEY) return make_int4(-1, -1, -1, -1);
        slot = (slot + 1) & mask;
    }
}


// ------------------------------------------------------------------------------------------------
// LEAF GRID.  ncu of the probe loop (profiles/r01_fwd_v7_summary.txt, r01_umma_notes.md section 6): 54 % issue active and the
// same ~100 G probes/s on a 2 MB table as on a 32 MB one -- the build is bound by the INSTRUCTIONS of 81-125 independent
// hash probes per row (hash, 64-bit compare, divergent linear-probe chains), not by where the slots come from.  A voxel's
// kernel neighbourhood is spatially compact, so the input set is also stored as a sparse grid of 4 x 4 x 4 (x 1 in t) leaves:
// a small hash table of leaf keys (8 bytes each) and, per table slot, the dense 64-entry array of the rows of the leaf's
// voxels.  Per output row the 3 x 3 x 3 (x 3) or 5 x 5 x 5 neighbourhood touches at most 2 x 2 x 2 (x 3) = 24 leaves: ONE
// warp iteration of parallel leaf probes, then every kernel offset is a shuffle + an indexed 4-byte load (no hash, no key
// compare, no chain) and the loads of a row fall into the few 256-byte leaf arrays it touches.
// Index space: coordinates of a strided level are multiples of step[d]; leaf coordinate = floor(c / step) >> lsh[d].
struct LgGeom { int step[4]; int lsh[4]; int mb[4]; int ncand; };
// Every key sits within LG_MAX_PROBE slots of its home slot, or the build raises the grid's overflow word and the map
// builders ignore the grid (plain voxel-table probes): lookups are bounded and never wrong, whatever the input.
#define LG_MAX_PROBE 64

__global__ void k_leafgrid_insert(const int32_t* __restrict__ coords, int64_t n, int ncol, LgGeom g,
                                  unsigned long long* __restrict__ keys, int32_t* __restrict__ rows, uint64_t mask,
                                  unsigned int* __restrict__ ok_word) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t* c = coords + i * ncol;
    int blk[4] = {0, 0, 0, 0}, loc = 0, bit = 0;
    for (int d = 0; d < ncol - 1; ++d) {
        const int idx = floor_div(c[1 + d], g.step[d]);
        blk[d] = idx >> g.lsh[d];
        loc |= (idx & ((1 << g.lsh[d]) - 1)) << bit;
        bit += g.lsh[d];
    }
    const uint64_t key = pack_key(c[0], blk[0], blk[1], blk[2], blk[3]);
    uint64_t slot = hash64(key) & mask;
    for (int step = 0; step < LG_MAX_PROBE; ++step) {
        unsigned long long* kp = keys + slot;
        const unsigned long long cur = *reinterpret_cast<volatile unsigned long long*>(kp);
        bool mine = cur == key;
        if (!mine && cur == INSMOS_EMPTY_KEY) {
            const unsigned long long prev = atomicCAS(kp, INSMOS_EMPTY_KEY, (unsigned long long)key);
            mine = prev == INSMOS_EMPTY_KEY || prev == key;
        }
        if (mine) { rows[slot * 64 + loc] = (int32_t)i; return; }
        slot = (slot + 1) & mask;
    }
    *ok_word = 0u;                                               // table too crowded for bounded probing: grid unusable
}
__device__ __forceinline__ int lg_find(const unsigned long long* __restrict__ keys, uint64_t mask, uint64_t key) {
    uint64_t slot = hash64(key) & mask;
    for (int step = 0; step < LG_MAX_PROBE; ++step) {
        const unsigned long long k = __ldg(keys + slot);
        if (k == key) return (int)slot;
        if (k == INSMOS_EMPTY_KEY) return -1;
        slot = (slot + 1) & mask;
    }
    return -1;                                                   // every stored key is within LG_MAX_PROBE of its home slot
}
extern "C" int64_t insmos_leafgrid_capacity(int64_t n) {
    // leaves <= voxels; LiDAR surfaces put ~8 voxels in a 4x4x4 leaf, so a table of >= n/2 slots runs at a load of ~0.25.
    // Inputs with more leaves than that (isolated voxels) overflow the bounded probing and fall back to the voxel table.
    int64_t cap = 1024;
    while (2 * cap < n) cap <<= 1;
    return cap;
}
extern "C" int64_t insmos_leafgrid_bytes(int64_t cap) { return cap <= 0 ? 0 : cap * 8 + cap * 64 * 4 + 16; }
static int lg_geometry(int ncol, const int32_t* step, LgGeom& g) {
    for (int d = 0; d < 4; ++d) { g.step[d] = 1; g.lsh[d] = 0; g.mb[d] = 1; }
    const int ndim = ncol - 1;
    if (ndim < 1 || ndim > 4 || !step) return INSMOS_ERR_INVALID_ARG;
    for (int d = 0; d < ndim; ++d) {
        if (step[d] < 1) return INSMOS_ERR_INVALID_ARG;
        g.step[d] = step[d];
        g.lsh[d] = d < 3 ? 2 : 0;                                  // 4 x 4 x 4 leaves in space, 1 in time
    }
    return INSMOS_OK;
}
extern "C" int insmos_leafgrid_build(const int32_t* coords, int64_t n, int32_t ncol, const int32_t* step,
                                     void* grid, int64_t cap, void* stream) {
    if (!grid || cap <= 0 || (cap & (cap - 1)) || cap > (1ll << 30) || (n > 0 && !coords) || n < 0 || ncol < 2 || ncol > 5)
        return INSMOS_ERR_INVALID_ARG;
    LgGeom g;
    const int rc = lg_geometry(ncol, step, g);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    INSMOS_CHECK_CUDA(cudaMemsetAsync(grid, 0xFF, (size_t)insmos_leafgrid_bytes(cap), st));   // keys = empty, rows = -1
    if (n > 0) {
        unsigned long long* keys = reinterpret_cast<unsigned long long*>(grid);
        int32_t* rows = reinterpret_cast<int32_t*>(keys + cap);
        unsigned int* ok_word = reinterpret_cast<unsigned int*>(rows + cap * 64);      // 0xffffffff after the memset = usable
        k_leafgrid_insert<<<(unsigned)ceil_div64(n, 256), 256, 0, st>>>(coords, n, ncol, g, keys, rows, (uint64_t)(cap - 1), ok_word);
        INSMOS_CHECK_LAUNCH("k_leafgrid_insert");
    }
    return INSMOS_OK;
}

__global__ void __launch_bounds__(RB_THREADS)
k_rulebook_tiles(const int32_t* __restrict__ out_coords, int64_t n_out,
                 const insmos_slot_t* __restrict__ table, uint64_t mask,
                 insmos_mapspec_t spec, int TM,
                 uint16_t* __restrict__ seg, uint32_t* __restrict__ entries,
                 unsigned long long* pair_count, const int32_t* __restrict__ parent,
                 const XBlockSlot* __restrict__ xtable, uint64_t xmask, int xstep,
                 const unsigned long long* __restrict__ lgkeys, const int32_t* __restrict__ lgrows, uint64_t lgmask, LgGeom lg,
                 const int32_t* __restrict__ first_row) {
    extern __shared__ __align__(16) int smem[];
    if (first_row && (int64_t)(blockIdx.x + 1) * TM <= (int64_t)__ldg(first_row)) return;     // tile not needed (dead-row elimination)
    const int K = spec.K, ncol = spec.ncol, ndim = spec.ndim;
    int* nbr = smem;                                            // [TM*K] in-row or -1
    int64_t* delta = reinterpret_cast<int64_t*>(nbr + ((TM * K + 1) & ~1));   // [K] packed key delta of offset k
    int64_t* rbase = delta + K;                                 // [TM] packed base key of the row (or -1)
    int* kd = reinterpret_cast<int*>(rbase + TM);               // [K*4] per-offset digits/terms (slow path)
    int* tc = kd + K * 4;                                       // [TM*5] coordinates of the tile's rows
    int* hist = tc + TM * 5;                                    // [K+1]
    int* kpk = hist + K + 1;                                    // [K] kernel digits packed one byte per dimension (leaf-grid path)
    int* rowblk = kpk + K;                                      // [TM*4] leaf coordinate of the row's neighbourhood corner
    unsigned* rowoff = reinterpret_cast<unsigned*>(rowblk + TM * 4);   // [TM] offset of the corner inside its leaf, one byte per dim
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = RB_THREADS / 32;
    const int64_t tile = blockIdx.x;
    const int64_t row0 = tile * TM;
    const bool fast = (spec.mode == 0) && spec.q[0] == 1 && spec.q[1] == 1 && spec.q[2] == 1 && spec.q[3] == 1;

    for (int i = tid; i < TM * ncol; i += RB_THREADS) {
        const int64_t gi = row0 * ncol + i;
        tc[(i / ncol) * 5 + (i % ncol)] = (gi < n_out * ncol) ? out_coords[gi] : 0;
    }
    for (int k = tid; k < K; k += RB_THREADS) {
        int rem = k;
        int dig[4] = {0, 0, 0, 0};
        if (spec.first_fastest) { for (int d = 0; d < ndim; ++d) { dig[d] = rem % spec.ksize[d]; rem /= spec.ksize[d]; } }
        else { for (int d = ndim - 1; d >= 0; --d) { dig[d] = rem % spec.ksize[d]; rem /= spec.ksize[d]; } }
        int term[4];
        for (int d = 0; d < 4; ++d) {
            term[d] = (d < ndim) ? (spec.mode == 0 ? dig[d] * spec.e[d] : dig[d]) : 0;
            kd[k * 4 + d] = (d < ndim && spec.mode == 0) ? spec.b[d] + term[d] : term[d];
        }
        delta[k] = packed_delta(term[0], term[1], term[2], term[3]);
        kpk[k] = dig[0] | (dig[1] << 8) | (dig[2] << 16) | (dig[3] << 24);
        hist[k] = 0;
    }
    __syncthreads();
    if (fast) {                                                  // base key of every row: pack(c*a + b)
        for (int r = tid; r < TM; r += RB_THREADS) {
            int64_t key = -1;
            if (row0 + r < n_out) {
                const int* c = tc + r * 5;
                int cb[4] = {0, 0, 0, 0};
                bool ok = true;
                for (int d = 0; d < 4; ++d)
                    if (d < ndim) {
                        cb[d] = c[1 + d] * spec.a[d] + spec.b[d];
                        const int lim = (d == 3) ? 128 - 16 : 32768 - RB_GUARD;
                        ok = ok && (cb[d] > -lim) && (cb[d] < lim);
                    }
                if (ok && (unsigned)c[0] < 128u) key = (int64_t)pack_key(c[0], cb[0], cb[1], cb[2], cb[3]);   // >= 0
                else key = -2;                                   // in range only with per-probe checks: slow path
            }
            rbase[r] = key;
        }
        __syncthreads();
    }

    // ---- phase A (leaf grid): one warp iteration of parallel leaf probes per row, then shuffle + indexed load per offset.
    // Leaves are 4 x 4 x 4 in the (up to three) spatial dimensions and 1 in time: the unpacking below is written for that
    // geometry (ncu of the first, generic version: the per-row set-up with four runtime divisions executed by every warp was
    // 21 % of all instructions -- it now runs once per row in a thread-per-row pre-pass, with shifts for power-of-two steps).
    const bool use_lg = lgkeys && fast && (*reinterpret_cast<const unsigned int*>(lgrows + (lgmask + 1) * 64) != 0u);
    if (use_lg) {
        for (int r = tid; r < TM; r += RB_THREADS) {
            // low corner of the row's neighbourhood in index space: leaf coordinate + offset inside the leaf (one byte per dim)
            const int* c = tc + r * 5;
            int blk[4] = {0, 0, 0, 0};
            unsigned off = 0u;
            bool lattice = (row0 + r) < n_out && rbase[r] >= 0;
#pragma unroll
            for (int d = 0; d < 4; ++d)
                if (d < ndim) {
                    const int lo = c[1 + d] * spec.a[d] + spec.b[d];
                    const int sh = lg.lsh[d];                            // here: log2(step) or -1
                    const int idx = sh >= 0 ? (lo >> sh) : floor_div(lo, lg.step[d]);
                    lattice = lattice && (idx * lg.step[d] == lo);     // off the input lattice: no neighbour can exist
                    if (d < 3) { blk[d] = idx >> 2; off |= (unsigned)(idx & 3) << (8 * d); }
                    else blk[d] = idx;
                }
            rowblk[r * 4 + 0] = blk[0]; rowblk[r * 4 + 1] = blk[1]; rowblk[r * 4 + 2] = blk[2]; rowblk[r * 4 + 3] = blk[3];
            rowoff[r] = lattice ? off : 0xffffffffu;
        }
        __syncthreads();
        int cj[4];                                               // this lane's candidate leaf (j0..j3) -- a function of the lane only
        { int t = lane; for (int d = 0; d < 4; ++d) { cj[d] = t % lg.mb[d]; t /= lg.mb[d]; } }
        const int s1 = lg.mb[0], s2 = lg.mb[0] * lg.mb[1], s3 = lg.mb[0] * lg.mb[1] * lg.mb[2];
        // a lane owns the offsets k = lane, lane + 32, ... (<= LG_KI of them) for every row of its warp: their packed digits and
        // bucket counts stay in registers (one shared-memory atomic per (warp, offset) at the end instead of one per pair),
        // and the indexed loads of a row are all issued before the first is consumed
        constexpr int LG_KI = 4;                                 // K <= 128
        unsigned kp[LG_KI];
        int cnt[LG_KI];
#pragma unroll
        for (int u = 0; u < LG_KI; ++u) { const int k = lane + 32 * u; kp[u] = k < K ? (unsigned)kpk[k] : 0u; cnt[u] = 0; }
        const int nki = (K + 31) >> 5;
        for (int r = warp; r < TM; r += nwarps) {
            const bool row_ok = (row0 + r) < n_out;
            const int* c = tc + r * 5;
            if (row_ok && rbase[r] == -2) {                          // near the edge of the packable range: checked path, voxel table
                for (int k = lane; k < K; k += 32) {
                    int ci[4] = {0, 0, 0, 0};
#pragma unroll
                    for (int d = 0; d < 4; ++d) if (d < ndim) ci[d] = c[1 + d] * spec.a[d] + kd[k * 4 + d];
                    int res = -1;
                    if (coord_in_range(c[0], ci[0], ci[1], ci[2], ci[3]))
                        res = table_find_row(table, mask, pack_key(c[0], ci[0], ci[1], ci[2], ci[3]));
                    nbr[r * K + k] = res;
                    if (res >= 0) atomicAdd(&hist[k], 1);
                }
                continue;
            }
            const unsigned off = rowoff[r];
            int mybase = -1;                                         // first element of this lane's candidate leaf in lgrows, or -1
            if (off != 0xffffffffu && lane < lg.ncand) {
                const int* b = rowblk + r * 4;
                const int slot = lg_find(lgkeys, lgmask, pack_key(c[0], b[0] + cj[0], b[1] + cj[1], b[2] + cj[2], b[3] + cj[3]));
                mybase = slot >= 0 ? slot << 6 : -1;
            }
            int res[LG_KI];
#pragma unroll
            for (int u = 0; u < LG_KI; ++u) {
                res[u] = -1;
                if (u < nki) {                                       // warp-uniform: all lanes take part in the shuffle
                    const unsigned v = off + kp[u];                  // byte d = offset in the leaf + kernel digit (< 64)
                    const unsigned t0 = v & 0x00030303u;
                    const unsigned loc = (t0 | (t0 >> 6) | (t0 >> 12)) & 0x3fu;
                    const unsigned jj = v >> 2;
                    const int cand = (int)(jj & 0x3fu) + (int)((jj >> 8) & 0x3fu) * s1 + (int)((jj >> 16) & 0x3fu) * s2 + (int)(v >> 24) * s3;
                    const int base = __shfl_sync(0xffffffffu, mybase, cand & 31);
                    if (base >= 0 && lane + 32 * u < K) res[u] = __ldg(lgrows + (unsigned)(base + (int)loc));
                }
            }
#pragma unroll
            for (int u = 0; u < LG_KI; ++u)
                if (u < nki && lane + 32 * u < K) {
                    nbr[r * K + lane + 32 * u] = res[u];
                    cnt[u] += res[u] >= 0 ? 1 : 0;
                }
        }
#pragma unroll
        for (int u = 0; u < LG_KI; ++u)
            if (cnt[u] > 0) atomicAdd(&hist[lane + 32 * u], cnt[u]);
    } else
    // ---- phase A (x-block table): lanes run over the offset GROUPS (all dimensions but x), warps over rows; a lane
    // resolves the ksize[0] consecutive x of its group with one or two 32-byte slot lookups.
    if (xtable && fast) {
        const int kx = spec.ksize[0], ng = K / kx;
        for (int r = warp; r < TM; r += nwarps) {
            const bool row_ok = (row0 + r) < n_out;
            const int64_t base = rbase[r];
            if (row_ok && base == -2) {                              // near the edge of the packable range: checked path, voxel table
                const int* c = tc + r * 5;
                for (int k = lane; k < K; k += 32) {
                    int ci[4] = {0, 0, 0, 0};
#pragma unroll
                    for (int d = 0; d < 4; ++d) if (d < ndim) ci[d] = c[1 + d] * spec.a[d] + kd[k * 4 + d];
                    int res = -1;
                    if (coord_in_range(c[0], ci[0], ci[1], ci[2], ci[3]))
                        res = table_find_row(table, mask, pack_key(c[0], ci[0], ci[1], ci[2], ci[3]));
                    nbr[r * K + k] = res;
                    if (res >= 0) atomicAdd(&hist[k], 1);
                }
                continue;
            }
            for (int g = lane; g < ng; g += 32) {
                int res[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) res[j] = -1;
                if (row_ok && base >= 0) {
                    const uint64_t key0 = (uint64_t)(base + delta[g * kx]);          // voxel key of the run's first x
                    const int x0 = (int)(key0 & 0xffffull) - 32768;
                    const int xi0 = floor_div(x0, xstep);
                    const uint64_t rest = key0 & ~0xffffull;
                    const int b_first = xi0 >> 2, b_last = (xi0 + kx - 1) >> 2;
                    for (int b = b_first; b <= b_last; ++b) {
                        const int4 rows = xblock_find(xtable, xmask, rest | (uint64_t)(uint32_t)(b + 32768));
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const int xi = xi0 + j;
                            if (j < kx && (xi >> 2) == b) {
                                const int sub = xi & 3;
                                res[j] = sub == 0 ? rows.x : sub == 1 ? rows.y : sub == 2 ? rows.z : rows.w;
                            }
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (j < kx) {
                        nbr[r * K + g * kx + j] = res[j];
                        if (res[j] >= 0) atomicAdd(&hist[g * kx + j], 1);
                    }
            }
        }
    } else
    // ---- phase A: probe.  Lanes run over offsets, warps over rows.
    for (int r = warp; r < TM; r += nwarps) {
        const bool row_ok = (row0 + r) < n_out;
        const int64_t base = fast ? rbase[r] : -2;
        for (int k = lane; k < K; k += 32) {
            int res = -1;
            if (row_ok) {
                if (base >= 0) {
                    res = table_find_row(table, mask, (uint64_t)(base + delta[k]));
                } else if (base == -2) {                         // checked path (also: strided/inverse/transposed maps)
                    const int* c = tc + r * 5;
                    int ci[4] = {0, 0, 0, 0};
                    bool ok = true;
                    if (spec.mode == 0) {
#pragma unroll
                        for (int d = 0; d < 4; ++d)
                            if (d < ndim) {
                                int v = c[1 + d] * spec.a[d] + kd[k * 4 + d];
                                const int q = spec.q[d];
                                if (q > 1) { if (v % q) ok = false; v /= q; }
                                ci[d] = v;
                            }
                    } else {
#pragma unroll
                        for (int d = 0; d < 4; ++d)
                            if (d < ndim) {
                                const int b0 = floor_div(c[1 + d], spec.up_q[d]) * spec.up_q[d];
                                if ((c[1 + d] - b0) / spec.up_ts[d] != kd[k * 4 + d]) ok = false;
                                ci[d] = b0;
                            }
                    }
                    if (ok && parent) res = __ldg(parent + row0 + r);        // transposed map: the one pair of a fine row is (k, its parent)
                    else if (ok && coord_in_range(c[0], ci[0], ci[1], ci[2], ci[3]))
                        res = table_find_row(table, mask, pack_key(c[0], ci[0], ci[1], ci[2], ci[3]));
                }
            }
            nbr[r * K + k] = res;
            if (res >= 0) atomicAdd(&hist[k], 1);               // bucket sizes (shared-memory integer atomics are native)
        }
    }
    __syncthreads();
    if (warp == 0) {                                  // exclusive scan of hist[0..K) -> hist, hist[K] = total
        int carry = 0;
        for (int k0 = 0; k0 < K; k0 += 32) {
            const int k = k0 + lane;
            const int v = (k < K) ? hist[k] : 0;
            int inc = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
            if (k < K) hist[k] = carry + inc - v;
            carry += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (lane == 0) {
            hist[K] = carry;
            if (pair_count && carry) atomicAdd(pair_count, (unsigned long long)carry);
        }
    }
    __syncthreads();
    uint16_t* tseg = seg + tile * (K + 1);
    for (int k = tid; k <= K; k += RB_THREADS) tseg[k] = (uint16_t)hist[k];

    // ---- phase C: compact each bucket in output-row order
    uint32_t* tent = entries + tile * (int64_t)TM * K;
    for (int k = warp; k < K; k += nwarps) {
        int pos = hist[k];
        for (int r0 = 0; r0 < TM; r0 += 32) {
            const int r = r0 + lane;
            const int v = (r < TM) ? nbr[r * K + k] : -1;
            const unsigned bal = __ballot_sync(0xffffffffu, v >= 0);
            if (v >= 0) tent[pos + __popc(bal & ((1u << lane) - 1u))] = ((uint32_t)r << INSMOS_ROW_BITS) | (uint32_t)v;
            pos += __popc(bal);
        }
    }
}

static int rulebook_build_impl(const int32_t* out_coords, int64_t n_out,
                               const insmos_slot_t* in_table, int64_t in_cap, const int32_t* parent,
                               const insmos_mapspec_t* spec, int32_t TM,
                               uint16_t* seg, uint32_t* entries, unsigned long long* pair_count, void* stream,
                               const void* xtable = nullptr, int64_t xcap = 0, int32_t xstep = 0,
                               const void* lgrid = nullptr, int64_t lgcap = 0, const int32_t* lgstep = nullptr,
                               const int32_t* first_row = nullptr) {
    if (!out_coords || !spec || !seg || !entries || n_out < 0) return INSMOS_ERR_INVALID_ARG;
    if (parent) {
        if (spec->mode != 1) return INSMOS_ERR_INVALID_ARG;
        in_cap = 1;                                                    // no table: every lookup is parent[row]
    } else {
        if (!in_table) return INSMOS_ERR_INVALID_ARG;
    }
    if (in_cap <= 0 || (in_cap & (in_cap - 1))) return INSMOS_ERR_INVALID_ARG;
    if (TM != 16 && TM != 32 && TM != 64 && TM != 128) return INSMOS_ERR_INVALID_ARG;
    if (spec->K <= 0 || (int64_t)TM * spec->K >= 65536 || spec->ndim < 1 || spec->ndim > 4 ||
        (spec->ncol != 4 && spec->ncol != 5) || spec->ndim > spec->ncol - 1)
        return INSMOS_ERR_INVALID_ARG;
    int kprod = 1;
    for (int d = 0; d < spec->ndim; ++d) {
        if (spec->ksize[d] < 1) return INSMOS_ERR_INVALID_ARG;
        if (spec->mode == 0 && spec->q[d] < 1) return INSMOS_ERR_INVALID_ARG;
        if (spec->mode == 0 && (spec->ksize[d] - 1) * (spec->e[d] < 0 ? -spec->e[d] : spec->e[d]) >= (d == 3 ? 16 : RB_GUARD / 2))
            return INSMOS_ERR_UNSUPPORTED;
        if (spec->mode == 1 && (spec->up_q[d] < 1 || spec->up_ts[d] < 1)) return INSMOS_ERR_INVALID_ARG;
        kprod *= spec->ksize[d];
    }
    if (kprod != spec->K) return INSMOS_ERR_INVALID_ARG;
    if (n_out == 0) return INSMOS_OK;
    const size_t smem = sizeof(int) * (((size_t)TM * spec->K + 1) & ~(size_t)1) + sizeof(int64_t) * ((size_t)spec->K + TM) +
                        sizeof(int) * ((size_t)spec->K * 4 + (size_t)TM * 5 + spec->K + 1 + spec->K + (size_t)TM * 5);
    LgGeom lg;
    for (int d = 0; d < 4; ++d) { lg.step[d] = 1; lg.lsh[d] = 0; lg.mb[d] = 1; }
    lg.ncand = 0;
    const unsigned long long* lgkeys = nullptr;
    const int32_t* lgrows = nullptr;
    if (lgrid) {
        if (lgcap <= 0 || (lgcap & (lgcap - 1)) || spec->mode != 0) return INSMOS_ERR_INVALID_ARG;
        const int rc = lg_geometry(spec->ncol, lgstep, lg);
        if (rc) return rc;
        int ncand = 1;
        for (int d = 0; d < spec->ndim; ++d) {
            if (spec->q[d] != 1 || spec->e[d] != lg.step[d]) return INSMOS_ERR_UNSUPPORTED;      // digits must walk the input lattice
            if (spec->ksize[d] > 60) return INSMOS_ERR_UNSUPPORTED;
            // leaves a run of ksize indices can touch: 4-wide leaves in space (worst start = 3), 1-wide in time
            lg.mb[d] = d < 3 ? ((spec->ksize[d] + 2) >> 2) + 1 : spec->ksize[d];
            ncand *= lg.mb[d];
            int sh = -1;                                               // the kernel reads lsh[] as log2(step) (or -1: divide)
            for (int b = 0; b < 16; ++b) if (lg.step[d] == (1 << b)) sh = b;
            lg.lsh[d] = sh;
        }
        if (ncand > 32 || spec->K > 128) return INSMOS_ERR_UNSUPPORTED;
        lg.ncand = ncand;
        lgkeys = reinterpret_cast<const unsigned long long*>(lgrid);
        lgrows = reinterpret_cast<const int32_t*>(lgkeys + lgcap);
    }
    if (smem > 220 * 1024) return INSMOS_ERR_UNSUPPORTED;
    static thread_local insmos_smem_cfg_t configured;
    INSMOS_CHECK_CUDA(insmos_ensure_smem(k_rulebook_tiles, smem, configured));
    const int64_t n_tiles = ceil_div64(n_out, TM);
    k_rulebook_tiles<<<(unsigned)n_tiles, RB_THREADS, smem, (cudaStream_t)stream>>>(
        out_coords, n_out, in_table, (uint64_t)(in_cap - 1), *spec, TM, seg, entries, pair_count, parent,
        reinterpret_cast<const XBlockSlot*>(xtable), (uint64_t)(xcap > 0 ? xcap - 1 : 0), xstep,
        lgkeys, lgrows, (uint64_t)(lgcap > 0 ? lgcap - 1 : 0), lg, first_row);
    INSMOS_CHECK_LAUNCH("k_rulebook_tiles");
    return INSMOS_OK;
}

extern "C" int insmos_rulebook_build(const int32_t* out_coords, int64_t n_out,
                                     const insmos_slot_t* in_table, int64_t in_cap,
                                     const insmos_mapspec_t* spec, int32_t TM,
                                     uint16_t* seg, uint32_t* entries, unsigned long long* pair_count,
                                     int32_t* counters, void* stream) {
    (void)counters;
    return rulebook_build_impl(out_coords, n_out, in_table, in_cap, nullptr, spec, TM, seg, entries, pair_count, stream);
}

// one block per tile, one thread per fine row: offset digit from the coordinates, parent row from the inverse map,
// counting sort by offset inside the tile (bucket order = output-row order, as k_rulebook_tiles produces)
__global__ void __launch_bounds__(128)
k_rulebook_up(const int32_t* __restrict__ coords, int64_t n_out, const int32_t* __restrict__ parent,
              insmos_mapspec_t spec, int TM, uint16_t* __restrict__ seg, uint32_t* __restrict__ entries,
              unsigned long long* pair_count) {
    __shared__ int kk[128];
    __shared__ int hist[130];
    const int K = spec.K, ncol = spec.ncol, ndim = spec.ndim;
    const int tid = threadIdx.x;
    const int64_t tile = blockIdx.x, row0 = tile * TM;
    for (int k = tid; k <= K; k += 128) hist[k] = 0;
    __syncthreads();
    int k = -1, par = 0;
    if (tid < TM && row0 + tid < n_out) {
        const int32_t* c = coords + (row0 + tid) * ncol;
        int dig[4] = {0, 0, 0, 0};
        bool ok = true;
        for (int d = 0; d < ndim; ++d) {
            const int v = c[1 + d];
            const int b0 = floor_div(v, spec.up_q[d]) * spec.up_q[d];
            dig[d] = (v - b0) / spec.up_ts[d];
            ok = ok && dig[d] < spec.ksize[d];
        }
        if (ok) {
            k = 0;
            if (spec.first_fastest) { for (int d = ndim - 1; d >= 0; --d) k = k * spec.ksize[d] + dig[d]; }
            else { for (int d = 0; d < ndim; ++d) k = k * spec.ksize[d] + dig[d]; }
            par = __ldg(parent + row0 + tid);
            atomicAdd(&hist[k], 1);
        }
    }
    if (tid < TM) kk[tid] = k;
    __syncthreads();
    if (tid == 0) {
        int run = 0;
        for (int j = 0; j < K; ++j) { const int v = hist[j]; hist[j] = run; run += v; }
        hist[K] = run;
        if (pair_count && run) atomicAdd(pair_count, (unsigned long long)run);
    }
    __syncthreads();
    uint16_t* tseg = seg + tile * (K + 1);
    for (int j = tid; j <= K; j += 128) tseg[j] = (uint16_t)hist[j];
    if (k >= 0) {
        int rank = 0;
        for (int r = 0; r < tid; ++r) rank += (kk[r] == k);
        entries[tile * (int64_t)TM * K + hist[k] + rank] = ((uint32_t)tid << INSMOS_ROW_BITS) | (uint32_t)par;
    }
}

// Transposed (up-sampling) map without hash probes: a fine row's only pair is (offset of the row inside its coarse
// cell, its parent row), and the parent row is the inverse map that insmos_unique_coords(q) returned when the coarse
// coordinate set was made.  Same rule-book layout / ordering as insmos_rulebook_build with a mode-1 spec.
extern "C" int insmos_rulebook_build_up(const int32_t* fine_coords, int64_t n_fine, const int32_t* parent,
                                        const insmos_mapspec_t* spec, int32_t TM,
                                        uint16_t* seg, uint32_t* entries, unsigned long long* pair_count, void* stream) {
    if (!parent) return INSMOS_ERR_INVALID_ARG;
    if (fine_coords && spec && seg && entries && n_fine > 0 && spec->mode == 1 && spec->K >= 1 && spec->K <= 128 &&
        (TM == 16 || TM == 32 || TM == 64 || TM == 128) && spec->ndim >= 1 && spec->ndim <= 4 &&
        (spec->ncol == 4 || spec->ncol == 5) && spec->ndim <= spec->ncol - 1) {
        int kprod = 1;
        bool ok = true;
        for (int d = 0; d < spec->ndim; ++d) { ok = ok && spec->ksize[d] >= 1 && spec->up_q[d] >= 1 && spec->up_ts[d] >= 1; kprod *= spec->ksize[d]; }
        if (ok && kprod == spec->K) {
            k_rulebook_up<<<(unsigned)ceil_div64(n_fine, TM), 128, 0, (cudaStream_t)stream>>>(
                fine_coords, n_fine, parent, *spec, TM, seg, entries, pair_count);
            INSMOS_CHECK_LAUNCH("k_rulebook_up");
            return INSMOS_OK;
        }
    }
    return rulebook_build_impl(fine_coords, n_fine, nullptr, 1, parent, spec, TM, seg, entries, pair_count, stream);
}

// Same map as insmos_rulebook_build for an affine cube spec whose x runs are contiguous (mode 0, q = 1, a[0] = 1,
// e[0] = xstep, first dimension fastest, 3 <= ksize[0] <= 8), probing the x-block table of the input set.  Bit-identical
// output; the voxel table is still needed for rows at the edge of the packable coordinate range.
extern "C" int insmos_rulebook_build_xb(const int32_t* out_coords, int64_t n_out,
                                        const insmos_slot_t* in_table, int64_t in_cap,
                                        const void* xtable, int64_t xcap, int32_t xstep,
                                        const insmos_mapspec_t* spec, int32_t TM,
                                        uint16_t* seg, uint32_t* entries, unsigned long long* pair_count, void* stream) {
    if (!xtable || !spec || xcap <= 0 || (xcap & (xcap - 1)) || xstep < 1) return INSMOS_ERR_INVALID_ARG;
    if (spec->mode != 0 || !spec->first_fastest || spec->a[0] != 1 || spec->e[0] != xstep || spec->ksize[0] < 3 || spec->ksize[0] > 8)
        return INSMOS_ERR_UNSUPPORTED;
    for (int d = 0; d < spec->ndim; ++d) if (spec->q[d] != 1) return INSMOS_ERR_UNSUPPORTED;
    return rulebook_build_impl(out_coords, n_out, in_table, in_cap, nullptr, spec, TM, seg, entries, pair_count, stream,
                               xtable, xcap, xstep);
}

// Same map as insmos_rulebook_build for an affine spec (mode 0, q = 1) whose kernel digits walk the input lattice
// (e[d] == step[d]), probing the LEAF GRID of the input set (insmos_leafgrid_build).  Bit-identical output; the voxel table is
// still needed for rows at the edge of the packable coordinate range.
extern "C" int insmos_rulebook_build_lg(const int32_t* out_coords, int64_t n_out,
                                        const insmos_slot_t* in_table, int64_t in_cap,
                                        const void* grid, int64_t grid_cap, const int32_t* step,
                                        const insmos_mapspec_t* spec, int32_t TM,
                                        uint16_t* seg, uint32_t* entries, unsigned long long* pair_count, void* stream) {
    if (!grid || !spec || !step) return INSMOS_ERR_INVALID_ARG;
    return rulebook_build_impl(out_coords, n_out, in_table, in_cap, nullptr, spec, TM, seg, entries, pair_count, stream,
                               nullptr, 0, 0, grid, grid_cap, step);
}

// insmos_rulebook_build_lg for the output tiles holding rows >= *first_row only (dead-row elimination)
extern "C" int insmos_rulebook_build_lg_from(const int32_t* out_coords, int64_t n_out,
                                             const insmos_slot_t* in_table, int64_t in_cap,
                                             const void* grid, int64_t grid_cap, const int32_t* step,
                                             const insmos_mapspec_t* spec, int32_t TM,
                                             uint16_t* seg, uint32_t* entries, unsigned long long* pair_count,
                                             const int32_t* first_row, void* stream) {
    if (!grid || !spec || !step) return INSMOS_ERR_INVALID_ARG;
    return rulebook_build_impl(out_coords, n_out, in_table, in_cap, nullptr, spec, TM, seg, entries, pair_count, stream,
                               nullptr, 0, 0, grid, grid_cap, step, first_row);
}
