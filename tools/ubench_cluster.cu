// micro-benchmark: cost of cluster barriers and of the engine's all-reduce primitives on a 16-CTA cluster
#include <cstdio>
#include "../open-simulator_b200/csrc/simon_kernel.cuh"

extern "C" __global__ void k_sync(long long *out, int iters, int mode) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cg::cluster_group cluster = cg::this_cluster();
    SkSmem S;
    const uint32_t NW = (blockDim.x + 31) >> 5;
    sk_carve(S, smem_raw, blockDim.x, 2, 4, 64, cluster.num_blocks());
    for (uint32_t q = threadIdx.x; q < (B_N32 + 2 + 4) * blockDim.x; q += blockDim.x) S.a32[q] = (int32_t)q;
    for (uint32_t q = threadIdx.x; q < C_N8 * blockDim.x; q += blockDim.x) S.a8[q] = (uint8_t)q;
    SkRed R{&S, &cluster, cluster.block_rank(), cluster.num_blocks(), 0};
    sk_red_init(R);
    long long t0 = clock64();
    unsigned long long acc = threadIdx.x;
    for (int i = 0; i < iters; i++) {
        if (mode == 0) cluster.sync();
        else if (mode == 1) { unsigned long long v[5] = {acc, acc + 1, acc + 2, acc + 3, acc & 1}; const int op[5] = {OP_MINU, OP_MAXU, OP_MINU, OP_MAXU, OP_OR}; sk_allreduce<5>(R, v, op); acc += v[1]; }
        else if (mode == 2) { unsigned long long v[16]; int op[16]; for (int q = 0; q < 16; q++) { v[q] = acc + q; } const int opc[16] = {0,0,2,2,2,1,1,2,1,2,3,3,3,3,3,3}; sk_allreduce<16>(R, v, opc); acc += v[2]; }
        else if (mode == 3) { uint32_t who; unsigned long long k = sk_argmax(R, ((acc & 0xffff) << 24) | (0xFFFFFFu - (cluster.block_rank() * blockDim.x + threadIdx.x)), cluster.num_blocks() * blockDim.x, blockDim.x, who); acc += k + sk_wpay(S, who, 0) + sk_wpay(S, who, 8); }
        else if (mode == 4) __syncthreads();
        else if (mode == 5) { __threadfence(); cluster.sync(); }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = (t1 - t0) / iters; out[1] = (long long)acc; }
}

int main() {
    long long *d; cudaMalloc(&d, 64);
    const char *names[6] = {"cluster.sync", "allreduce<5>", "allreduce<16>", "argmax", "__syncthreads", "threadfence+cluster.sync"};
    for (int cs : {1, 4, 16}) for (int tpb : {128, 256}) for (int mode = 0; mode < 6; mode++) {
        size_t smem = sk_smem_bytes(tpb, 2, 4, 64, cs);
        cudaFuncSetAttribute(k_sync, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(k_sync, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(cs); cfg.blockDim = dim3(tpb); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        cudaError_t e = cudaLaunchKernelEx(&cfg, k_sync, d, 2000, mode);
        cudaError_t e2 = cudaDeviceSynchronize();
        long long h[4] = {0, 0, 0, 0};
        cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
        printf("cs=%2d tpb=%4d %-28s %6lld cycles  acc=%lld (%s %s)\n", cs, tpb, names[mode], h[0], h[1], cudaGetErrorString(e), cudaGetErrorString(e2));
    }
    return 0;
}
