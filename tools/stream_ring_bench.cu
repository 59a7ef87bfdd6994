// Micro-benchmark (measurement tool, not product): what is the fastest a single launch can
// stream N bytes of HBM into the SMs on this GPU, (a) with the bulk-copy (TMA 1-D) ring the
// probe kernel uses, without any compute, and (b) with plain 16-byte loads?  Prints the
// time per launch (CUDA events around K back-to-back launches on one stream and round-robin on two,
// buffers cycled so no launch hits L2).  Used to separate "ring design limit" from "probe compute".
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/stream_ring_bench tools/stream_ring_bench.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// (a) ring of S stages of `chunk` bytes; CTA b owns chunks b, b+G, ...; `split` bulk copies per chunk.
// Every warp touches one word of the chunk (so the data is really consumed), the last warp out refills.
__global__ void __launch_bounds__(1024, 1)
ring_kernel(const uint8_t* __restrict__ src, uint32_t n_chunks, uint32_t chunk, uint32_t S, uint32_t split,
            uint32_t* __restrict__ sink) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint32_t* done = reinterpret_cast<uint32_t*>(smem + 128);
    uint8_t* stages = smem + 256;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5, G = gridDim.x;
    if (tid == 0) {
        for (uint32_t s = 0; s < S; ++s) { mbar_init(&full[s], 1); done[s] = 0; }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint32_t my = n_chunks > blockIdx.x ? (n_chunks - blockIdx.x + G - 1) / G : 0;
    auto fill = [&](uint32_t s, uint32_t it) {
        const uint8_t* p = src + static_cast<size_t>(blockIdx.x + it * G) * chunk;
        mbar_expect_tx(&full[s], chunk);
        const uint32_t part = chunk / split;
        for (uint32_t j = 0; j < split; ++j) bulk_g2s(stages + static_cast<size_t>(s) * chunk + j * part, p + j * part, part, &full[s]);
    };
    if (warp == 0 && lane < S && lane < my) fill(lane, lane);
    uint32_t s = 0, ph = 0, acc = 0;
    for (uint32_t it = 0; it < my; ++it) {
        mbar_wait(&full[s], ph);
        acc += *reinterpret_cast<const uint32_t*>(stages + static_cast<size_t>(s) * chunk + ((tid * 64u) % chunk));
        __syncwarp();
        if (lane == 0) {
            const uint32_t old = atomicAdd(&done[s], 1u);
            if (old == nw - 1) {
                __threadfence_block();
                done[s] = 0;
                if (it + S < my) { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); fill(s, it + S); }
            }
        }
        if (++s == S) { s = 0; ph ^= 1u; }
    }
    if (acc == 0x12345678u) sink[0] = acc;
}

// (b) plain vector loads, grid-stride over 16-byte words, UNROLL independent loads in flight per thread
template <int UNROLL>
__global__ void __launch_bounds__(1024, 1)
ldg_kernel(const uint4* __restrict__ src, size_t n16, uint32_t* __restrict__ sink) {
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    uint32_t acc = 0;
    for (; i + (UNROLL - 1) * stride < n16; i += UNROLL * stride) {
        uint4 v[UNROLL];
#pragma unroll
        for (int j = 0; j < UNROLL; ++j) v[j] = __ldcs(src + i + j * stride);
#pragma unroll
        for (int j = 0; j < UNROLL; ++j) acc += v[j].x ^ v[j].y ^ v[j].z ^ v[j].w;
    }
    for (; i < n16; i += stride) { uint4 v = __ldcs(src + i); acc += v.x ^ v.w; }
    if (acc == 0x12345678u) sink[0] = acc;
}

int main(int argc, char** argv) {
    const size_t total = argc > 1 ? strtoull(argv[1], nullptr, 10) : 70482888ull;  // bytes per launch
    const int K = 200, NBUF = 8;
    int sms = 0, smem_max = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    CK(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, 0));
    const size_t buf_bytes = (total + (1u << 20)) & ~size_t(255);
    std::vector<uint8_t*> bufs(NBUF);
    for (auto& b : bufs) { CK(cudaMalloc(&b, buf_bytes)); CK(cudaMemset(b, 1, buf_bytes)); }
    uint32_t* sink; CK(cudaMalloc(&sink, 4));
    cudaStream_t st[3]; for (auto& s : st) CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    cudaEvent_t e0, e1, ef[2]; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (auto& e : ef) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    CK(cudaFuncSetAttribute(ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max));
    printf("device SMs %d, smem opt-in %d, bytes per launch %zu (ideal at 6540 GB/s: %.2f us)\n", sms, smem_max, total,
           total / 6540.2e3);

    auto timeit = [&](const char* name, int n_streams, auto&& launch) {
        for (int i = 0; i < 20; ++i) launch(i % NBUF, st[0]);
        CK(cudaStreamSynchronize(st[0]));
        CK(cudaEventRecord(e0, st[0]));
        if (n_streams == 1) {
            for (int i = 0; i < K; ++i) launch(i % NBUF, st[0]);
        } else {  // fork two worker streams from st[0], issue round-robin, join
            CK(cudaEventRecord(ef[0], st[0]));
            for (int j = 0; j < 2; ++j) CK(cudaStreamWaitEvent(st[1 + j], ef[0], 0));
            for (int i = 0; i < K; ++i) launch(i % NBUF, st[1 + (i & 1)]);
            for (int j = 0; j < 2; ++j) { CK(cudaEventRecord(ef[j], st[1 + j])); CK(cudaStreamWaitEvent(st[0], ef[j], 0)); }
        }
        CK(cudaEventRecord(e1, st[0]));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        const double us = ms * 1e3 / K;
        printf("%-44s streams %d: %7.2f us/launch  %7.1f GB/s\n", name, n_streams, us, total / us / 1e3);
    };

    for (int ns = 1; ns <= 2; ++ns) {
        timeit("ldg16 unroll 8, 148x1024", ns, [&](int b, cudaStream_t s) {
            ldg_kernel<8><<<sms, 1024, 0, s>>>(reinterpret_cast<const uint4*>(bufs[b]), total / 16, sink); });
        timeit("ldg16 unroll 4, 296x512", ns, [&](int b, cudaStream_t s) {
            ldg_kernel<4><<<2 * sms, 512, 0, s>>>(reinterpret_cast<const uint4*>(bufs[b]), total / 16, sink); });
        timeit("ldg16 unroll 4, 1184x256", ns, [&](int b, cudaStream_t s) {
            ldg_kernel<4><<<8 * sms, 256, 0, s>>>(reinterpret_cast<const uint4*>(bufs[b]), total / 16, sink); });
        const uint32_t chunks[] = {70656, 35328, 17664, 8832, 4416};
        for (uint32_t chunk : chunks) {
            const uint32_t n_chunks = static_cast<uint32_t>(total / chunk);
            const uint32_t budget = smem_max - 256;
            const uint32_t maxS = budget / chunk;
            for (uint32_t S : {maxS > 16 ? 16u : maxS, 2u}) {
                if (S > maxS || S == 0) continue;
                for (uint32_t split : {1u, 4u}) {
                    for (uint32_t warps : {32u, 4u}) {
                        char name[128];
                        snprintf(name, sizeof name, "ring chunk %u S %u split %u warps %u", chunk, S, split, warps);
                        timeit(name, ns, [&](int b, cudaStream_t s) {
                            ring_kernel<<<sms, warps * 32, 256 + S * chunk, s>>>(bufs[b], n_chunks, chunk, S, split, sink); });
                    }
                }
            }
        }
    }
    CK(cudaGetLastError());
    return 0;
}
