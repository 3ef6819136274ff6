// device_prims.cuh -- small device-wide primitives shared by lcb_device.cu and graph_device.cu: exclusive scan (u32),
// stable LSD radix sort of a permutation by 8-bit digits of a key word, gather, iota.  Include INSIDE an anonymous
// namespace: every translation unit gets its own copies.
#pragma once

__global__ void k_iota(unsigned *list, unsigned n, unsigned first = 0, unsigned stride = 1)
{
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) list[i] = first + i * stride;
}

// ---- exclusive scan (u32), three small kernels -------------------------------------------------------
constexpr int kScanTile = 2048;
__global__ void __launch_bounds__(256) k_scan_tiles(const unsigned *in, unsigned *out, unsigned *tile_sum, size_t n)
{
    __shared__ unsigned s[256];
    const size_t base = (size_t)blockIdx.x * kScanTile + (size_t)threadIdx.x * 8;
    unsigned v[8], sum = 0;
    for (int q = 0; q < 8; q++) {
        v[q] = base + q < n ? in[base + q] : 0;
        sum += v[q];
    }
    s[threadIdx.x] = sum;
    __syncthreads();
    for (int d = 1; d < 256; d <<= 1) {
        unsigned x = threadIdx.x >= d ? s[threadIdx.x - d] : 0;
        __syncthreads();
        s[threadIdx.x] += x;
        __syncthreads();
    }
    unsigned run = s[threadIdx.x] - sum;
    for (int q = 0; q < 8; q++) {
        if (base + q < n) out[base + q] = run;
        run += v[q];
    }
    if (threadIdx.x == 255) tile_sum[blockIdx.x] = s[255];
}
__global__ void k_scan_sums(unsigned *tile_sum, unsigned tiles, unsigned *total)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        unsigned run = 0;
        for (unsigned i = 0; i < tiles; i++) {
            unsigned x = tile_sum[i];
            tile_sum[i] = run;
            run += x;
        }
        *total = run;
    }
}
__global__ void k_scan_add(unsigned *out, const unsigned *tile_sum, size_t n)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] += tile_sum[i / kScanTile];
}

// ---- stable LSD radix sort of a permutation by 8-bit digits of a key word ---------------------------------
constexpr int kSortTile = 2048; // elements per block
template <typename K>
__global__ void __launch_bounds__(256) k_radix_hist(const unsigned *perm, const K *key, int shift, unsigned n, unsigned *hist, int flip)
{
    __shared__ unsigned h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const unsigned base = blockIdx.x * kSortTile;
    for (int q = 0; q < kSortTile / 256; q++) {
        unsigned i = base + q * 256 + threadIdx.x;
        if (i < n) {
            unsigned d = (unsigned)(key[perm[i]] >> shift) & 255u;
            if (flip) d = 255u - d;
            atomicAdd(&h[d], 1u);
        }
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * gridDim.x + blockIdx.x] = h[threadIdx.x]; // digit-major
}
template <typename K>
__global__ void __launch_bounds__(256) k_radix_scatter(const unsigned *perm_in, unsigned *perm_out, const K *key, int shift,
                                                         unsigned n, const unsigned *offs, int flip)
{
    __shared__ unsigned run[256];
    __shared__ unsigned wc[8][256];
    run[threadIdx.x] = offs[(size_t)threadIdx.x * gridDim.x + blockIdx.x];
    for (int w = 0; w < 8; w++) wc[w][threadIdx.x] = 0;
    __syncthreads();
    const unsigned base = blockIdx.x * kSortTile;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int q = 0; q < kSortTile / 256; q++) {
        unsigned i = base + q * 256 + threadIdx.x;
        bool live = i < n;
        unsigned p = live ? perm_in[i] : 0;
        unsigned d = live ? ((unsigned)(key[p] >> shift) & 255u) : 256u + (unsigned)lane; // dead lanes match nobody
        if (live && flip) d = 255u - d;
        unsigned peers = __match_any_sync(0xFFFFFFFFu, d);
        unsigned before = __popc(peers & ((1u << lane) - 1));
        if (live && before == 0) wc[warp][d] = __popc(peers);
        __syncthreads();
        if (live) {
            unsigned pre = 0;
            for (int w = 0; w < warp; w++) pre += wc[w][d];
            perm_out[run[d] + pre + before] = p;
        }
        __syncthreads();
        {
            unsigned tot = 0;
            for (int w = 0; w < 8; w++) {
                tot += wc[w][threadIdx.x];
                wc[w][threadIdx.x] = 0;
            }
            run[threadIdx.x] += tot;
        }
        __syncthreads();
    }
}
template <typename T>
__global__ void k_gather(const unsigned *perm, const T *in, T *out, unsigned n)
{
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[perm[i]];
}
__global__ void k_is_uniform_digit(const unsigned *hist, unsigned blocks, unsigned n, unsigned *flag)
{
    // one CTA per digit (launch with 256 CTAs of 256 threads): if a single digit holds all n keys the pass is the identity
    __shared__ unsigned part[8];
    const unsigned d = blockIdx.x;
    unsigned tot = 0;
    for (unsigned b = threadIdx.x; b < blocks; b += blockDim.x) tot += hist[(size_t)d * blocks + b];
    tot = __reduce_add_sync(0xFFFFFFFFu, tot);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = tot;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned all = 0;
        for (unsigned w = 0; w < (blockDim.x >> 5); w++) all += part[w];
        if (all == n) *flag = 1;
    }
}

