// Experiment helper (not product): does __viaddmin_s16x2(a, b, 0x7FFF7FFF) add the halves with 16-bit wrap-around?
#include <cstdio>
#include <cstdint>
__global__ void k(unsigned* bad, unsigned* first)
{
    const uint32_t a = blockIdx.x * blockDim.x + threadIdx.x;        // 0..65535
    for (uint32_t b = 0; b < 65536; b += 1) {
        const uint32_t pa = a | ((a ^ 0x5A5Au) << 16), pb = b | ((b * 31u & 0xFFFFu) << 16);
        const uint32_t got = __viaddmin_s16x2(pa, pb, 0x7FFF7FFFu);
        const uint32_t want = ((pa + pb) & 0xFFFFu) | ((((pa >> 16) + (pb >> 16)) & 0xFFFFu) << 16);
        if (got != want) { if (atomicAdd(bad, 1u) == 0) { first[0] = pa; first[1] = pb; first[2] = got; first[3] = want; } }
    }
}
int main()
{
    unsigned *bad, *first; cudaMalloc(&bad, 4); cudaMalloc(&first, 16); cudaMemset(bad, 0, 4); cudaMemset(first, 0, 16);
    k<<<256, 256>>>(bad, first);
    unsigned hb = 0, hf[4];
    cudaMemcpy(&hb, bad, 4, cudaMemcpyDeviceToHost); cudaMemcpy(hf, first, 16, cudaMemcpyDeviceToHost);
    printf("viaddmin_s16x2 wrap test over 2^32 pairs: %u mismatches (first: a=%08x b=%08x got=%08x want=%08x) err=%s\n", hb, hf[0], hf[1], hf[2], hf[3], cudaGetErrorString(cudaGetLastError()));
    return 0;
}
