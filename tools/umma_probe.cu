// Hardware probe (not product code): does a K-major SWIZZLE_128B UMMA operand tolerate a start address that is
// shifted by whole 128-byte rows (not a multiple of the 1024-byte swizzle atom), and 8-row groups whose
// stride (SBO) is not a multiple of 1024?  This decides whether the 3x3 convolutions can reuse ONE halo tile in
// shared memory for all nine filter taps.  Build: make -C tools ; run on a B200: tools/build/umma_probe
#include "../dl_vqa_b200/csrc/tc_common.cuh"
#include <vector>
#include <cstdlib>
#include <cmath>

using namespace tc;

constexpr int ROWS = 256, N = 64, K = 64;

__device__ __forceinline__ uint64_t desc_variant(uint32_t saddr, uint32_t sbo, int mode) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(sbo >> 4) << 32;
    d |= (uint64_t)1 << 46;
    if (mode == 1) d |= (uint64_t)((saddr >> 7) & 7u) << 49;      // base_offset = row phase inside the atom
    d |= (uint64_t)2 << 61;
    return d;
}

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap ta, const __grid_constant__ CUtensorMap tb, float* out,
             int shift_rows, int sbo, int mode) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sa = smem;                       // 256 rows x 128 B
    uint8_t* sb = smem + ROWS * 128;          // 64 rows x 128 B
    uint64_t* bar = reinterpret_cast<uint64_t*>(sb + N * 128);
    uint32_t* tbase = reinterpret_cast<uint32_t*>(bar + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(tbase, 64);
    tcgen05_fence_before(); __syncthreads(); tcgen05_fence_after();
    const uint32_t tm = *tbase;
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bar[0], ROWS * 128 + N * 128);
        tma_load_2d(sa, &ta, &bar[0], 0, 0);
        tma_load_2d(sb, &tb, &bar[0], 0, 0);
        mbar_wait(&bar[0], 0);
        tcgen05_fence_after();
        const uint32_t a_addr = smem_u32(sa) + shift_rows * 128, b_addr = smem_u32(sb);
        for (int k = 0; k < 4; ++k)
            umma_f16(tm, desc_variant(a_addr + k * 32, sbo, mode), desc_variant(b_addr + k * 32, 1024, 0),
                     idesc_bf16(128, N), k > 0);
        umma_commit(&bar[1]);
    }
    mbar_wait(&bar[1], 0);
    tcgen05_fence_after();
    for (int c0 = 0; c0 < N; c0 += 32) {
        float v[32];
        tmem_ld_32x32(tm + c0 + ((uint32_t)(warp * 32) << 16), v);
        for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * N + c0 + j] = v[j];
    }
    tcgen05_fence_before(); __syncthreads();
    if (warp == 0) { tcgen05_fence_after(); tmem_dealloc(tm, 64); }
}

static float bf(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

int main() {
    std::vector<bf16> ha(ROWS * K), hb(N * K);
    std::vector<float> fa(ROWS * K), fb(N * K);
    srand(1);
    for (int i = 0; i < ROWS * K; ++i) { fa[i] = bf((rand() % 17 - 8) / 8.f); ha[i] = __float2bfloat16_rn(fa[i]); }
    for (int i = 0; i < N * K; ++i) { fb[i] = bf((rand() % 13 - 6) / 4.f); hb[i] = __float2bfloat16_rn(fb[i]); }
    bf16 *da, *db; float* dout;
    cudaMalloc(&da, ha.size() * 2); cudaMalloc(&db, hb.size() * 2); cudaMalloc(&dout, 128 * N * 4);
    cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
    CUtensorMap ta, tb;
    { uint64_t d[2] = {K, ROWS}; uint64_t s[1] = {K * 2}; uint32_t b[2] = {64, ROWS}; if (make_tmap_bf16(&ta, da, 2, d, s, b)) return 2; }
    { uint64_t d[2] = {K, N}; uint64_t s[1] = {K * 2}; uint32_t b[2] = {64, N}; if (make_tmap_bf16(&tb, db, 2, d, s, b)) return 2; }
    const int smem = ROWS * 128 + N * 128 + 1024 + 64;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    std::vector<float> got(128 * N);
    const int sbos[3] = {1024, 1280, 2048};
    for (int si = 0; si < 3; ++si)
        for (int mode = 0; mode < 2; ++mode)
            for (int shift = 0; shift < 10; ++shift) {
                const int sbo = sbos[si];
                if (shift + 15 * (sbo / 128) + 8 > ROWS) continue;
                cudaMemset(dout, 0, 128 * N * 4);
                probe_kernel<<<1, 128, smem>>>(ta, tb, dout, shift, sbo, mode);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("sbo %d mode %d shift %d: CUDA error %s\n", sbo, mode, shift, cudaGetErrorString(e)); return 3; }
                cudaMemcpy(got.data(), dout, 128 * N * 4, cudaMemcpyDeviceToHost);
                double maxerr = 0; int bad = 0;
                for (int m = 0; m < 128; ++m) {
                    const int row = shift + (m / 8) * (sbo / 128) + (m % 8);
                    for (int n = 0; n < N; ++n) {
                        double ref = 0;
                        for (int k = 0; k < K; ++k) ref += (double)fa[row * K + k] * fb[n * K + k];
                        const double err = fabs(ref - got[m * N + n]);
                        if (err > 1e-3) ++bad;
                        if (err > maxerr) maxerr = err;
                    }
                }
                printf("sbo %4d  base_offset_mode %d  row_shift %d : %s (bad %d / %d, max err %.4f)\n", sbo, mode, shift,
                       bad ? "MISMATCH" : "ok", bad, 128 * N, maxerr);
            }
    return 0;
}
