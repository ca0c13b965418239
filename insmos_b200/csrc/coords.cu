// Coordinate ops: hashed unique in first-occurrence order, fused 4D quantise+unique (a1,a2),
// ME coordinate stride, spconv output-coordinate generation, capped 3D voxelisation + mean VFE (a6,a7).
//
// Canonical row order.  MinkowskiEngine / spconv GPU builds number voxels in hash order (non
// deterministic across runs); their CPU builds number them by first occurrence.  The oracle
// (oracle/) freezes first-occurrence order and these kernels reproduce it exactly without a sort:
//   1. every input i inserts its key and does atomicMin(slot.first, i)
//   2. input i is the "first" of its voxel iff slot.first == i; an order-preserving exclusive scan
//      of that flag over the inputs is the voxel's row id.
// All reads of the input are coalesced streams; the only random traffic is the 16-byte slot.
#include "common.cuh"

static thread_local char g_last_error[512] = "";
void insmos_set_last_error(const char* what, cudaError_t e) {
    snprintf(g_last_error, sizeof(g_last_error), "%s: %s", what, cudaGetErrorString(e));
}
extern "C" const char* insmos_last_error(void) { return g_last_error; }
unsigned long long g_insmos_launches = 0ull;
extern "C" uint64_t insmos_launch_count(void) { return __atomic_load_n(&g_insmos_launches, __ATOMIC_RELAXED); }
extern "C" const char* insmos_version(void) { return "insmos_b200 0.1 (sm_100a)"; }

extern "C" int64_t insmos_hash_capacity(int64_t n) {
    // n = number of INSERTIONS (points / candidate coordinates); the distinct keys are ~0.2-0.45 n on LiDAR input, so 2n slots
    // run at a load of 0.1-0.25 (<= 0.5 even if every insertion is distinct: linear probing always terminates).  Round 1 used
    // 4n: the 134 MB table of the 1.2 M-point cloud cost 45 us per forward just to clear, and the big kernel maps, the reason
    // for the very low load (a warp pays for its LONGEST probe chain), now go through the leaf grid instead.
    int64_t cap = 1024;
    while (cap < 2 * n) cap <<= 1;
    return cap;
}
extern "C" int64_t insmos_scan_scratch_bytes(int64_t n) {
    return (ceil_div64(n > 0 ? n : 1, INSMOS_SCAN_BLOCK) + 1) * (int64_t)sizeof(unsigned long long);
}

// ------------------------------------------------------------------------------------------------
__global__ void k_table_clear(insmos_slot_t* table, int64_t cap) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < cap) {
        int4 v; v.x = -1; v.y = -1; v.z = INT_MAX; v.w = -1;     // key = all ones, first = INT_MAX, row = -1
        reinterpret_cast<int4*>(table)[i] = v;
    }
}
extern "C" int insmos_table_clear(insmos_slot_t* table, int64_t cap, void* stream) {
    if (!table || cap <= 0 || (cap & (cap - 1))) return INSMOS_ERR_INVALID_ARG;
    k_table_clear<<<(unsigned)ceil_div64(cap, 256), 256, 0, (cudaStream_t)stream>>>(table, cap);
    INSMOS_CHECK_LAUNCH("table_clear");
    return INSMOS_OK;
}

__global__ void k_scan_blocksums(unsigned long long* blocksums, int64_t nblocks, int32_t* counters,
                                 int c_lo, int c_hi) {
    // single block; exclusive scan in place, chunk by chunk
    unsigned long long carry = 0;
    for (int64_t base = 0; base < nblocks; base += blockDim.x) {
        int64_t i = base + threadIdx.x;
        unsigned long long v = (i < nblocks) ? blocksums[i] : 0ull;
        unsigned long long tot;
        unsigned long long ex = block_exclusive_scan_u64(v, &tot);
        if (i < nblocks) blocksums[i] = carry + ex;
        carry += tot;
    }
    if (threadIdx.x == 0) {
        if (c_lo >= 0) counters[c_lo] = (int32_t)(carry & 0xffffffffull);
        if (c_hi >= 0) counters[c_hi] = (int32_t)(carry >> 32);
    }
}

// ------------------------------------------------------------------------------------------------
// coordinate sources
struct Src4D {           // raw points (x,y,z,intensity,t), fp32 true division then floor (motionnet.py:25-28)
    const float* pts; int stride; float qx, qy, qz, qt;
    static constexpr int NCOL = 5;
    __device__ __forceinline__ bool load(int64_t i, int& b, int& c0, int& c1, int& c2, int& c3, bool& aux) const {
        const float* p = pts + i * stride;
        const float fx = __fdiv_rn(p[0], qx), fy = __fdiv_rn(p[1], qy), fz = __fdiv_rn(p[2], qz), ft = __fdiv_rn(p[4], qt);
        aux = (ft == 0.0f);                                  // motionnet.py:42 tests the divided, unfloored time
        b = 0; c0 = (int)floorf(fx); c1 = (int)floorf(fy); c2 = (int)floorf(fz); c3 = (int)floorf(ft);
        return true;
    }
    __device__ __forceinline__ void store(int32_t* out, int64_t row, int b, int c0, int c1, int c2, int c3) const {
        int32_t* o = out + row * 5; o[0] = b; o[1] = c0; o[2] = c1; o[3] = c2; o[4] = c3;
    }
};
struct SrcInt {          // int32 rows (b, c0..), optional floor to multiples of q (ME stride)
    const int32_t* coords; int ncol; int q0, q1, q2, q3;
    __device__ __forceinline__ bool load(int64_t i, int& b, int& c0, int& c1, int& c2, int& c3, bool& aux) const {
        const int32_t* p = coords + i * ncol;
        aux = false;
        b = p[0]; c0 = p[1]; c1 = p[2]; c2 = p[3]; c3 = (ncol > 4) ? p[4] : 0;
        if (q0 > 1) c0 = floor_div(c0, q0) * q0;
        if (q1 > 1) c1 = floor_div(c1, q1) * q1;
        if (q2 > 1) c2 = floor_div(c2, q2) * q2;
        if (q3 > 1) c3 = floor_div(c3, q3) * q3;
        return true;
    }
    __device__ __forceinline__ void store(int32_t* out, int64_t row, int b, int c0, int c1, int c2, int c3) const {
        int32_t* o = out + row * ncol; o[0] = b; o[1] = c0; o[2] = c1; o[3] = c2; if (ncol > 4) o[4] = c3;
    }
};
struct SrcSpOut {        // virtual element v = i*K + k of a strided spconv: candidate output coordinate
    const int32_t* coords; int K, kz, ky, kx, sz, sy, sx, pz, py, px, oz, oy, ox;
    __device__ __forceinline__ bool load(int64_t v, int& b, int& c0, int& c1, int& c2, int& c3, bool& aux) const {
        const int64_t i = v / K; const int k = (int)(v - i * K);
        const int dx = k % kx, dy = (k / kx) % ky, dz = k / (kx * ky);
        const int32_t* p = coords + i * 4;
        aux = false; b = p[0]; c3 = 0;
        int z = p[1] + pz - dz, y = p[2] + py - dy, x = p[3] + px - dx;
        if ((z % sz) | (y % sy) | (x % sx)) return false;
        z /= sz; y /= sy; x /= sx;
        if (z < 0 || z >= oz || y < 0 || y >= oy || x < 0 || x >= ox) return false;
        c0 = z; c1 = y; c2 = x;
        return true;
    }
    __device__ __forceinline__ void store(int32_t* out, int64_t row, int b, int c0, int c1, int c2, int c3) const {
        int32_t* o = out + row * 4; o[0] = b; o[1] = c0; o[2] = c1; o[3] = c2;
    }
};
struct SrcVox3D {        // points (x,y,z,...) clipped to a range, voxel = floor((p-min)/v), stored (0,z,y,x)
    const float* pts; int C; float x0, y0, z0, vx, vy, vz; int gx, gy, gz;
    __device__ __forceinline__ bool load(int64_t i, int& b, int& c0, int& c1, int& c2, int& c3, bool& aux) const {
        const float* p = pts + i * C;
        const int ix = (int)floorf(__fdiv_rn(__fsub_rn(p[0], x0), vx));
        const int iy = (int)floorf(__fdiv_rn(__fsub_rn(p[1], y0), vy));
        const int iz = (int)floorf(__fdiv_rn(__fsub_rn(p[2], z0), vz));
        aux = false; b = 0; c3 = 0;
        if (ix < 0 || ix >= gx || iy < 0 || iy >= gy || iz < 0 || iz >= gz) return false;
        c0 = iz; c1 = iy; c2 = ix;
        return true;
    }
    __device__ __forceinline__ void store(int32_t* out, int64_t row, int b, int c0, int c1, int c2, int c3) const {
        int32_t* o = out + row * 4; o[0] = b; o[1] = c0; o[2] = c1; o[3] = c2;
    }
};

#define SP_AUX_BIT 30
#define SP_SLOT_MASK ((1 << SP_AUX_BIT) - 1)

template <class Src>
__global__ void __launch_bounds__(INSMOS_SCAN_BLOCK)
k_insert(Src src, int64_t n, insmos_slot_t* table, uint64_t mask, int32_t* slot_of, int32_t* counters) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int b, c0, c1, c2, c3; bool aux;
    int32_t sp = -1;
    if (src.load(i, b, c0, c1, c2, c3, aux)) {
        if (!coord_in_range(b, c0, c1, c2, c3)) {
            atomicOr(&counters[INSMOS_CNT_ERR], INSMOS_DEVERR_COORD_RANGE);
        } else {
            const int64_t slot = table_insert(table, mask, pack_key(b, c0, c1, c2, c3));
            atomicMin(&table[slot].first, (int)i);
            sp = (int32_t)slot | ((aux ? 1 : 0) << SP_AUX_BIT);
        }
    }
    slot_of[i] = sp;
}

__device__ __forceinline__ unsigned long long first_aux_flags(const insmos_slot_t* table, int32_t sp, int64_t i) {
    if (sp < 0) return 0ull;
    const unsigned long long is_first = (table[sp & SP_SLOT_MASK].first == (int)i) ? 1ull : 0ull;
    const unsigned long long is_aux = (unsigned long long)((sp >> SP_AUX_BIT) & 1);
    return is_first | (is_aux << 32);
}

__global__ void __launch_bounds__(INSMOS_SCAN_BLOCK)
k_flag_blocksum(const insmos_slot_t* table, const int32_t* slot_of, int64_t n, unsigned long long* blocksums) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long v = (i < n) ? first_aux_flags(table, slot_of[i], i) : 0ull;
    unsigned long long tot;
    block_exclusive_scan_u64(v, &tot);
    if (threadIdx.x == 0) blocksums[blockIdx.x] = tot;
}

template <class Src>
__global__ void __launch_bounds__(INSMOS_SCAN_BLOCK)
k_finalize(Src src, int64_t n, insmos_slot_t* table, const int32_t* slot_of, const unsigned long long* blocksums,
           int32_t* out_coords, int32_t* aux_index, int32_t max_rows) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int32_t sp = (i < n) ? slot_of[i] : -1;
    const unsigned long long v = (i < n) ? first_aux_flags(table, sp, i) : 0ull;
    const unsigned long long ex = block_exclusive_scan_u64(v, nullptr) + blocksums[blockIdx.x];
    if (i >= n || sp < 0) return;
    if (v & 1ull) {
        const int32_t row = (int32_t)(ex & 0xffffffffull);
        if (row < max_rows) {
            table[sp & SP_SLOT_MASK].row = row;
            int b, c0, c1, c2, c3; bool aux;
            src.load(i, b, c0, c1, c2, c3, aux);
            src.store(out_coords, row, b, c0, c1, c2, c3);
        }
    }
    if ((v >> 32) && aux_index) aux_index[(int32_t)(ex >> 32)] = (int32_t)i;
}

__global__ void k_inverse(const insmos_slot_t* table, const int32_t* slot_of, int64_t n, int32_t* inverse) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t sp = slot_of[i];
    inverse[i] = (sp < 0) ? -1 : table[sp & SP_SLOT_MASK].row;
}

__global__ void k_clamp_rows(int32_t* counters, int32_t max_rows) {
    const int32_t total = counters[INSMOS_CNT_ROWS];
    counters[INSMOS_CNT_TOTAL] = total;
    counters[INSMOS_CNT_ROWS] = total < max_rows ? total : max_rows;
}

template <class Src>
static int run_unique(const Src& src, int64_t n, insmos_slot_t* table, int64_t cap, int32_t* slot_of,
                      int32_t* out_coords, int32_t* inverse, int32_t* aux_index, int32_t* counters,
                      void* scratch, int32_t max_rows, cudaStream_t st) {
    if (!table || cap <= 0 || (cap & (cap - 1)) || cap > (1ll << SP_AUX_BIT) || !slot_of || !counters || !scratch || n < 0 ||
        n >= (1ll << 31))
        return INSMOS_ERR_INVALID_ARG;
    INSMOS_CHECK_CUDA(cudaMemsetAsync(counters, 0, sizeof(int32_t) * INSMOS_NUM_COUNTERS, st));
    if (n == 0) return INSMOS_OK;
    const unsigned nb = (unsigned)ceil_div64(n, INSMOS_SCAN_BLOCK);
    unsigned long long* blocksums = (unsigned long long*)scratch;
    k_insert<Src><<<nb, INSMOS_SCAN_BLOCK, 0, st>>>(src, n, table, (uint64_t)(cap - 1), slot_of, counters);
    INSMOS_CHECK_LAUNCH("k_insert");
    k_flag_blocksum<<<nb, INSMOS_SCAN_BLOCK, 0, st>>>(table, slot_of, n, blocksums);
    INSMOS_CHECK_LAUNCH("k_flag_blocksum");
    k_scan_blocksums<<<1, INSMOS_SCAN_BLOCK, 0, st>>>(blocksums, nb, counters, INSMOS_CNT_ROWS, INSMOS_CNT_AUX);
    INSMOS_CHECK_LAUNCH("k_scan_blocksums");
    k_finalize<Src><<<nb, INSMOS_SCAN_BLOCK, 0, st>>>(src, n, table, slot_of, blocksums, out_coords, aux_index, max_rows);
    INSMOS_CHECK_LAUNCH("k_finalize");
    if (max_rows != INT_MAX) {
        k_clamp_rows<<<1, 1, 0, st>>>(counters, max_rows);
        INSMOS_CHECK_LAUNCH("k_clamp_rows");
    }
    if (inverse) {
        k_inverse<<<(unsigned)ceil_div64(n, 256), 256, 0, st>>>(table, slot_of, n, inverse);
        INSMOS_CHECK_LAUNCH("k_inverse");
    }
    return INSMOS_OK;
}

extern "C" int insmos_voxelize4d(const float* points, int64_t n, int32_t point_stride, const float* quant,
                                 insmos_slot_t* table, int64_t cap, int32_t* slot_of_point,
                                 int32_t* out_coords, int32_t* inverse, int32_t* cur_index,
                                 int32_t* counters, void* scratch, void* stream) {
    if ((n > 0 && !points) || point_stride < 5 || !quant || !out_coords || !inverse) return INSMOS_ERR_INVALID_ARG;
    Src4D src{points, point_stride, quant[0], quant[1], quant[2], quant[3]};
    return run_unique(src, n, table, cap, slot_of_point, out_coords, inverse, cur_index, counters, scratch, INT_MAX,
                      (cudaStream_t)stream);
}

extern "C" int insmos_unique_coords(const int32_t* coords, int64_t n, int32_t ncol, const int32_t* q,
                                    insmos_slot_t* table, int64_t cap, int32_t* slot_of_point,
                                    int32_t* out_coords, int32_t* inverse,
                                    int32_t* counters, void* scratch, void* stream) {
    if ((n > 0 && !coords) || (ncol != 4 && ncol != 5) || !out_coords) return INSMOS_ERR_INVALID_ARG;
    SrcInt src{coords, ncol, 1, 1, 1, 1};
    if (q) { src.q0 = q[0]; src.q1 = q[1]; src.q2 = q[2]; src.q3 = (ncol > 4) ? q[3] : 1; }
    if (src.q0 < 1 || src.q1 < 1 || src.q2 < 1 || src.q3 < 1) return INSMOS_ERR_INVALID_ARG;
    return run_unique(src, n, table, cap, slot_of_point, out_coords, inverse, nullptr, counters, scratch, INT_MAX,
                      (cudaStream_t)stream);
}

extern "C" int64_t insmos_spconv_out_scratch_bytes(int64_t n, int32_t K) {
    const int64_t nv = n * K;
    return (insmos_scan_scratch_bytes(nv) + 255) / 256 * 256 + (nv > 0 ? nv : 1) * (int64_t)sizeof(int32_t);
}

extern "C" int insmos_spconv_out_coords(const int32_t* in_coords, int64_t n,
                                        const int32_t* ksize, const int32_t* stride, const int32_t* pad,
                                        const int32_t* out_shape,
                                        insmos_slot_t* table, int64_t cap,
                                        int32_t* out_coords, int32_t* counters, void* scratch, void* stream) {
    if ((n > 0 && !in_coords) || !ksize || !stride || !pad || !out_shape || !out_coords) return INSMOS_ERR_INVALID_ARG;
    const int K = ksize[0] * ksize[1] * ksize[2];
    if (K <= 0 || stride[0] < 1 || stride[1] < 1 || stride[2] < 1) return INSMOS_ERR_INVALID_ARG;
    SrcSpOut src{in_coords, K, ksize[0], ksize[1], ksize[2], stride[0], stride[1], stride[2],
                 pad[0], pad[1], pad[2], out_shape[0], out_shape[1], out_shape[2]};
    // slot_of scratch for the n*K virtual elements lives behind the block sums
    const int64_t nv = n * K;
    char* base = (char*)scratch;
    const int64_t off = (insmos_scan_scratch_bytes(nv) + 255) / 256 * 256;
    int32_t* slot_of = (int32_t*)(base + off);
    return run_unique(src, nv, table, cap, slot_of, out_coords, nullptr, nullptr, counters, scratch, INT_MAX,
                      (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------
// 3D voxelisation with caps + ids + mean (spconv PointToVoxel.generate_voxel_with_id + MeanVFE)
__global__ void k_v3d_init(int32_t* cnt, int32_t* best, int32_t max_voxels, int32_t mp) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < max_voxels) cnt[i] = 0;
    if (i < (int64_t)max_voxels * mp) best[i] = INT_MAX;
}

// every point joins its voxel; the mp smallest point indices per voxel are kept, sorted, by a
// cascade of atomicMin (slot s keeps the minimum of everything that reaches it and forwards the
// maximum), which is order-independent.
__global__ void k_v3d_assign(const insmos_slot_t* table, const int32_t* slot_of, int64_t n, int32_t mp,
                             int32_t* pc_voxel_id, int32_t* cnt, int32_t* best) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t sp = slot_of[i];
    const int32_t row = (sp < 0) ? -1 : table[sp & SP_SLOT_MASK].row;
    pc_voxel_id[i] = row;
    if (row < 0) return;
    atomicAdd(&cnt[row], 1);
    int v = (int)i;
    for (int s = 0; s < mp; ++s) {
        const int old = atomicMin(&best[(int64_t)row * mp + s], v);
        if (old == INT_MAX) break;
        v = old > v ? old : v;
    }
}

__global__ void k_v3d_reduce(const float* pts, int32_t C, int32_t mp, const int32_t* counters,
                             const int32_t* cnt, const int32_t* best,
                             int32_t* num_points, float* voxels, float* mean) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int32_t nrows = counters[INSMOS_CNT_ROWS];
    const int64_t r = idx / C; const int c = (int)(idx - r * C);
    if (r >= nrows) return;
    int num = cnt[r]; num = num < mp ? num : mp;
    float acc = 0.0f;
    for (int s = 0; s < mp; ++s) {
        float v = 0.0f;
        if (s < num) v = pts[(int64_t)best[r * mp + s] * C + c];
        if (voxels) voxels[(r * mp + s) * C + c] = v;
        acc = __fadd_rn(acc, v);                                  // slot order, as sum(dim=1)
    }
    const float denom = (float)(num < 1 ? 1 : num);                // clamp_min(num,1)
    mean[r * C + c] = __fdiv_rn(acc, denom);
    if (c == 0) num_points[r] = num;
}

extern "C" int insmos_voxelize3d(const float* points, int64_t n, int32_t C,
                                 const float* range, const float* vsize, const int32_t* grid,
                                 int32_t max_voxels, int32_t max_points,
                                 insmos_slot_t* table, int64_t cap, int32_t* slot_of_point,
                                 int32_t* coords, int32_t* num_points, float* voxels, float* mean,
                                 int32_t* pc_voxel_id, int32_t* work,
                                 int32_t* counters, void* scratch, void* stream) {
    if ((n > 0 && !points) || C < 3 || !range || !vsize || !grid || max_voxels <= 0 || max_points <= 0 || !coords || !num_points ||
        !mean || !pc_voxel_id || !work)
        return INSMOS_ERR_INVALID_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    SrcVox3D src{points, C, range[0], range[1], range[2], vsize[0], vsize[1], vsize[2], grid[0], grid[1], grid[2]};
    int rc = run_unique(src, n, table, cap, slot_of_point, coords, nullptr, nullptr, counters, scratch, max_voxels, st);
    if (rc != INSMOS_OK) return rc;
    if (n == 0) return INSMOS_OK;
    int32_t* cnt = work; int32_t* best = work + max_voxels;
    k_v3d_init<<<(unsigned)ceil_div64((int64_t)max_voxels * max_points, 256), 256, 0, st>>>(cnt, best, max_voxels, max_points);
    INSMOS_CHECK_LAUNCH("k_v3d_init");
    k_v3d_assign<<<(unsigned)ceil_div64(n, 256), 256, 0, st>>>(table, slot_of_point, n, max_points, pc_voxel_id, cnt, best);
    INSMOS_CHECK_LAUNCH("k_v3d_assign");
    const int64_t upper = (n < max_voxels ? n : (int64_t)max_voxels) * C;
    k_v3d_reduce<<<(unsigned)ceil_div64(upper, 256), 256, 0, st>>>(points, C, max_points, counters, cnt, best,
                                                                   num_points, voxels, mean);
    INSMOS_CHECK_LAUNCH("k_v3d_reduce");
    return INSMOS_OK;
}

// ------------------------------------------------------------------------------------------------
// Dead-row elimination for the 4D MotionNet decoder: starts[j] = smallest row index whose time
// coordinate is >= -j (j = 0..15), n when there is none.  Class of a row = clamp(-t, 0, 16); per-block minima of the row
// index per class in shared memory, one global atomicMin per class and block, then a 1-thread prefix minimum
// (a row of class c serves every threshold j >= c).
#define TRS_CLASSES 17
__global__ void k_time_row_starts(const int32_t* __restrict__ coords, int64_t n, int ncol, int tcol, int32_t* __restrict__ cls_min) {
    __shared__ int smin[TRS_CLASSES];
    if (threadIdx.x < TRS_CLASSES) smin[threadIdx.x] = INT_MAX;
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const int t = coords[i * ncol + tcol];
        const int c = t >= 0 ? 0 : (t < -16 ? 16 : -t);
        atomicMin(&smin[c], (int)i);
    }
    __syncthreads();
    if (threadIdx.x < TRS_CLASSES && smin[threadIdx.x] != INT_MAX) atomicMin(&cls_min[threadIdx.x], smin[threadIdx.x]);
}
__global__ void k_time_row_prefix(const int32_t* __restrict__ cls_min, int32_t n, int32_t* __restrict__ starts) {
    int m = n;
    for (int j = 0; j < 16; ++j) { m = min(m, cls_min[j]); starts[j] = m; }
}

extern "C" int insmos_time_row_starts(const int32_t* coords, int64_t n, int32_t ncol, int32_t tcol, int32_t* starts, void* stream) {
    if (!starts || n < 0 || n > INT_MAX || ncol <= 0 || tcol < 0 || tcol >= ncol || (n > 0 && !coords)) return INSMOS_ERR_INVALID_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    int32_t* cls_min = starts + 16;                                      // scratch: starts holds 16 + 17 ints
    INSMOS_CHECK_CUDA(cudaMemsetAsync(cls_min, 0x7f, sizeof(int32_t) * TRS_CLASSES, st));       // 0x7f7f7f7f > any row index
    if (n > 0) {
        k_time_row_starts<<<(unsigned)ceil_div64(n, 256), 256, 0, st>>>(coords, n, ncol, tcol, cls_min);
        INSMOS_CHECK_LAUNCH("k_time_row_starts");
    }
    k_time_row_prefix<<<1, 1, 0, st>>>(cls_min, (int32_t)n, starts);
    INSMOS_CHECK_LAUNCH("k_time_row_prefix");
    return INSMOS_OK;
}
