// Device-wide exclusive scan and a stable LSD radix sort of (key, value) pairs -- the
// "GPU radix sort" stage of the Morton LBVH build.  Hand-written (no CUB/Thrust): 8-bit
// digits, per-tile histograms, one scan of the digit-major histogram table, and a scatter
// whose in-tile ranking uses __match_any_sync so equal digits keep their input order
// (stability is what makes "ties by primitive index" hold for equal Morton codes).
#include "ftn_scene.h"

namespace ftn {

// ---- exclusive scan ---------------------------------------------------------------------------------
static constexpr int SCAN_THREADS = 256;
static constexpr int SCAN_ITEMS = 4;
static constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_tiles(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, uint32_t* __restrict__ tile_sums, size_t n) {
    __shared__ uint32_t warp_sums[SCAN_THREADS / 32];
    const size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    uint32_t sum = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) { v[i] = (base + i < n) ? in[base + i] : 0u; sum += v[i]; }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = (lane < SCAN_THREADS / 32) ? warp_sums[lane] : 0u;
        uint32_t wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += y; }
        if (lane < SCAN_THREADS / 32) warp_sums[lane] = wi - w;   // exclusive
        if (lane == SCAN_THREADS / 32 - 1 && tile_sums) tile_sums[blockIdx.x] = wi;
    }
    __syncthreads();
    uint32_t run = warp_sums[warp] + (incl - sum);
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) { if (base + i < n) out[base + i] = run; run += v[i]; }
}

__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_add_offsets(uint32_t* __restrict__ out, const uint32_t* __restrict__ tile_offsets, size_t n) {
    const size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
    const uint32_t off = tile_offsets[blockIdx.x];
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) if (base + i < n) out[base + i] += off;
}

// scratch: scan_scratch_elems(n) uint32.  Enqueues only: no allocation, no synchronisation.
size_t scan_scratch_elems(size_t n) {
    size_t total = 0;
    while (n > (size_t)SCAN_TILE) { n = (n + SCAN_TILE - 1) / SCAN_TILE; total += (n + 63) & ~(size_t)63; }
    return total + 64;
}
int exclusive_scan_u32(const uint32_t* d_in, uint32_t* d_out, size_t n, uint32_t* d_scratch, cudaStream_t st) {
    if (n == 0) return FTN_OK;
    const size_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    if (tiles == 1) {
        k_scan_tiles<<<1, SCAN_THREADS, 0, st>>>(d_in, d_out, nullptr, n);
        FTN_LAUNCHED();
        return FTN_OK;
    }
    uint32_t* d_sums = d_scratch;
    k_scan_tiles<<<(unsigned)tiles, SCAN_THREADS, 0, st>>>(d_in, d_out, d_sums, n);
    FTN_LAUNCHED();
    FTN_TRY(exclusive_scan_u32(d_sums, d_sums, tiles, d_scratch + ((tiles + 63) & ~(size_t)63), st));
    k_scan_add_offsets<<<(unsigned)tiles, SCAN_THREADS, 0, st>>>(d_out, d_sums, n);
    FTN_LAUNCHED();
    return FTN_OK;
}

// ---- radix sort ---------------------------------------------------------------------------------------
static constexpr int RS_THREADS = 256;
static constexpr int RS_WARPS = RS_THREADS / 32;
static constexpr int RS_ITEMS = 8;                       // keys per thread
static constexpr int RS_TILE = RS_THREADS * RS_ITEMS;    // 2048 keys per block
static constexpr int RS_RADIX = 256;

// key i of a tile lives at tile*RS_TILE + warp*(32*RS_ITEMS) + round*32 + lane: each warp owns a
// contiguous chunk and walks it in order, so (warp, round, lane) order == input order.
__device__ __forceinline__ size_t rs_index(int tile, int warp, int round, int lane) {
    return (size_t)tile * RS_TILE + (size_t)warp * (32 * RS_ITEMS) + (size_t)round * 32 + lane;
}

__global__ void __launch_bounds__(RS_THREADS)
k_radix_hist(const uint32_t* __restrict__ keys, size_t n, int shift, uint32_t* __restrict__ tile_hist, uint32_t n_tiles) {
    __shared__ uint32_t h[RS_RADIX];
    h[threadIdx.x] = 0u;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        const size_t i = rs_index(blockIdx.x, warp, r, lane);
        if (i < n) atomicAdd(&h[(keys[i] >> shift) & 0xFFu], 1u);
    }
    __syncthreads();
    tile_hist[(size_t)threadIdx.x * n_tiles + blockIdx.x] = h[threadIdx.x];   // digit-major
}

__global__ void __launch_bounds__(RS_THREADS)
k_radix_scatter(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, size_t n, int shift,
                const uint32_t* __restrict__ tile_base, uint32_t n_tiles) {
    __shared__ uint32_t wcount[RS_WARPS][RS_RADIX];
    for (int i = threadIdx.x; i < RS_WARPS * RS_RADIX; i += RS_THREADS) (&wcount[0][0])[i] = 0u;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lt = (1u << lane) - 1u;
    uint32_t key[RS_ITEMS], val[RS_ITEMS], local[RS_ITEMS];
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        const size_t i = rs_index(blockIdx.x, warp, r, lane);
        const bool valid = i < n;
        key[r] = valid ? keys_in[i] : 0u;
        val[r] = valid ? vals_in[i] : 0u;
        const uint32_t digit = (key[r] >> shift) & 0xFFu;
        // invalid lanes get a value no digit can take, so they match nobody
        const uint32_t peers = __match_any_sync(0xffffffffu, valid ? digit : (0x100u | (uint32_t)lane));
        const uint32_t rank = __popc(peers & lt);
        uint32_t prev = 0u;
        if (valid) prev = wcount[warp][digit];
        __syncwarp();
        if (valid && rank == 0u) wcount[warp][digit] = prev + __popc(peers);
        __syncwarp();
        local[r] = prev + rank;
    }
    __syncthreads();
    {   // exclusive scan over the warps, one digit per thread
        const int d = threadIdx.x;
        uint32_t run = 0u;
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) { const uint32_t c = wcount[w][d]; wcount[w][d] = run; run += c; }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        const size_t i = rs_index(blockIdx.x, warp, r, lane);
        if (i < n) {
            const uint32_t digit = (key[r] >> shift) & 0xFFu;
            const size_t dst = (size_t)tile_base[(size_t)digit * n_tiles + blockIdx.x] + wcount[warp][digit] + local[r];
            keys_out[dst] = key[r];
            vals_out[dst] = val[r];
        }
    }
}

// scratch: radix_sort_scratch_bytes(n) bytes, 256-byte aligned.  Enqueues only: no allocation, no
// synchronisation.  The result lands in d_keys / d_vals.
static size_t rs_tiles(size_t n) { return (n + RS_TILE - 1) / RS_TILE; }
static size_t align256(size_t b) { return (b + 255) & ~(size_t)255; }
size_t radix_sort_scratch_bytes(size_t n) {
    const size_t hist = (size_t)RS_RADIX * rs_tiles(n);
    return 2 * align256(n * 4) + align256(hist * 4) + align256(scan_scratch_elems(hist) * 4);
}
int radix_sort_pairs(uint32_t* d_keys, uint32_t* d_vals, size_t n, int bits, void* d_scratch, cudaStream_t st) {
    if (n < 2) return FTN_OK;
    const uint32_t n_tiles = (uint32_t)rs_tiles(n);
    char* p = (char*)d_scratch;
    uint32_t* d_k2 = (uint32_t*)p; p += align256(n * 4);
    uint32_t* d_v2 = (uint32_t*)p; p += align256(n * 4);
    uint32_t* d_hist = (uint32_t*)p; p += align256((size_t)RS_RADIX * n_tiles * 4);
    uint32_t* d_scan = (uint32_t*)p;
    uint32_t *kin = d_keys, *vin = d_vals, *kout = d_k2, *vout = d_v2;
    int passes = (bits + 7) / 8;
    if (passes & 1) passes += 1;   // even number of passes: the result lands in the caller's buffers
    for (int pass = 0; pass < passes; ++pass) {
        const int shift = 8 * pass;
        if (shift >= 32) {   // padding pass: plain copy keeps the ping-pong parity
            FTN_CUDA(cudaMemcpyAsync(kout, kin, n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
            FTN_CUDA(cudaMemcpyAsync(vout, vin, n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
        } else {
            k_radix_hist<<<n_tiles, RS_THREADS, 0, st>>>(kin, n, shift, d_hist, n_tiles);
            FTN_LAUNCHED();
            FTN_TRY(exclusive_scan_u32(d_hist, d_hist, (size_t)RS_RADIX * n_tiles, d_scan, st));
            k_radix_scatter<<<n_tiles, RS_THREADS, 0, st>>>(kin, vin, kout, vout, n, shift, d_hist, n_tiles);
            FTN_LAUNCHED();
        }
        uint32_t* t;
        t = kin; kin = kout; kout = t;
        t = vin; vin = vout; vout = t;
    }
    return FTN_OK;
}

}  // namespace ftn
