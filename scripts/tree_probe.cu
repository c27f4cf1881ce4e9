// tree_probe.cu -- time of the shared-memory tree of msm_comb.cuh on one block, with and without team additions
#include <cstdio>
#include <vector>
#include "../halo2-prover_b200/csrc/msm_comb.cuh"
using namespace h2b;
template <bool TEAM>
__global__ void __launch_bounds__(256) tree(const XYZZ *in, XYZZ *out, long long *cyc) {
    extern __shared__ uint4 smem[];
    XYZZ *sh = reinterpret_cast<XYZZ *>(smem);
    const uint32_t tid = threadIdx.x, nt = blockDim.x;
    store_xyzz(&sh[tid], load_xyzz(&in[tid]));
    __syncthreads();
    long long t0 = clock64();
    int lvl = 0;
    for (uint32_t stride = nt >> 1; stride > 0; stride >>= 1, lvl++) {
        if (TEAM && stride * 4 <= nt) {
            const uint32_t team = tid >> 2;
            const bool active = team < stride;
            const uint32_t mask = __ballot_sync(0xffffffffu, active);
            if (active) {
                const XYZZ a = load_xyzz(&sh[team]);
                const XYZZ b = load_xyzz(&sh[team + stride]);
                const XYZZ c = xyzz_add_team4(a, b, tid & 3, mask);
                __syncwarp(mask);
                if ((tid & 3) == 0) store_xyzz(&sh[team], c);
            }
        } else if (tid < stride) {
            XYZZ a = load_xyzz(&sh[tid]);
            XYZZ b = load_xyzz(&sh[tid + stride]);
            xyzz_add(a, b);
            store_xyzz(&sh[tid], a);
        }
        __syncthreads();
        if (tid == 0) cyc[lvl] = clock64() - t0;
    }
    if (tid == 0) store_xyzz(out, load_xyzz(&sh[0]));
}
int main() {
    XYZZ *in, *out;
    long long *cyc, h[8];
    cudaMalloc(&in, 256 * sizeof(XYZZ));
    cudaMalloc(&out, sizeof(XYZZ));
    cudaMalloc(&cyc, 64);
    std::vector<uint32_t> host(256 * 32);
    for (size_t i = 0; i < host.size(); i++) host[i] = (uint32_t)(i * 2654435761u) & 0x0fffffffu;
    cudaMemcpy(in, host.data(), host.size() * 4, cudaMemcpyHostToDevice);
    for (int team = 0; team < 2; team++) {
        for (int rep = 0; rep < 2; rep++) {
            if (team) tree<true><<<1, 256, 256 * sizeof(XYZZ)>>>(in, out, cyc);
            else tree<false><<<1, 256, 256 * sizeof(XYZZ)>>>(in, out, cyc);
            cudaDeviceSynchronize();
        }
        cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost);
        printf("%s:", team ? "team additions" : "plain         ");
        for (int l = 0; l < 8; l++) printf(" L%d %lld", l, h[l] - (l ? h[l - 1] : 0));
        printf("  total %lld cycles\n", h[7]);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
