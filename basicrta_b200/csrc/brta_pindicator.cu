// Per-datum cluster membership counts from the stored label rows: SURVEY.md 8(f-1).
//
// basicrta/gibbs.py:264-268 walks, for every retained sample row j and every active component k
// of that row, over np.where(indicator[j] == k) and adds 1 to pindicator[i, cluster(j, k)] -- a
// double Python loop over ~1000 rows x active components with an O(N) scan each.  Here it is one
// pass over the label tensor: thread = four data, rows streamed with coalesced 32-bit loads, the
// (row, component) -> increment table of a row chunk in shared memory, byte counters packed in registers.
// Integer work, HBM-bound: n_rows * n_data bytes in, 4 * n_data * n_clusters bytes out.
#include <stdint.h>

#include "../../include/basicrta_b200.h"
#include "brta_host.h"

namespace brta {

constexpr int PI_THREADS = 256;
constexpr int PI_UNROLL = 8;            // rows (4 labels each) in flight per thread
constexpr int PI_SMEM_TABLE = 32768;    // bytes of shared memory for the increment table of a row chunk

// The counters of a thread are bytes packed into NW 32-bit registers (cluster c = byte c & 3 of word
// c >> 2), and the (row, component) -> cluster table holds, per entry, the NW words to ADD for a datum
// labelled with that component: 1 << 8(c & 3) in word c >> 2, or zeros for a pair that is not counted.
// One label then costs a byte load, a table load and NW integer adds.  A byte counts at most 255 rows, so
// a block handles chunks of <= 255 rows (blockIdx.y) and adds its unpacked counters to global memory.
template <int NW>
__global__ void __launch_bounds__(PI_THREADS)
pindicator_kernel(const uint8_t* __restrict__ indicator, int64_t row_stride, int n_rows, int n_data,
                  const int8_t* __restrict__ cluster_of, int ncomp, int n_clusters, int rows_per_chunk,
                  int32_t* __restrict__ counts)
{
    extern __shared__ __align__(16) uint32_t pi_table[];    // [rows][ncomp + 1][NW]; entry ncomp = zeros
    const int tid = threadIdx.x;
    const int i0 = (blockIdx.x * PI_THREADS + tid) * 4;    // this thread's four data
    const int row0 = blockIdx.y * rows_per_chunk;
    const int rows = min(rows_per_chunk, n_rows - row0);
    const int entries = ncomp + 1;
    for (int x = tid; x < rows * entries; x += PI_THREADS) {
        const int j = x / entries, k = x - j * entries;
        const int c = k < ncomp ? (int)cluster_of[(size_t)(row0 + j) * ncomp + k] : -1;
#pragma unroll
        for (int w = 0; w < NW; ++w)
            pi_table[(size_t)x * NW + w] = (c >= 0 && c < n_clusters && (c >> 2) == w) ? (1u << (8 * (c & 3))) : 0u;
    }
    __syncthreads();
    if (i0 >= n_data) return;
    uint32_t acc[4][NW];                                    // [datum of the thread][word]
#pragma unroll
    for (int d = 0; d < 4; ++d)
#pragma unroll
        for (int w = 0; w < NW; ++w) acc[d][w] = 0u;
    const unsigned last = (unsigned)ncomp;
    const uint32_t table_addr = (uint32_t)__cvta_generic_to_shared(pi_table);
    const uint32_t row_bytes = (uint32_t)entries * NW * 4u;
    // entry of (row j, label lab) by 32-bit shared-window address: one LEA + one LDS per label
    auto add = [&](int d, uint32_t row_addr, unsigned lab) {
        const uint32_t a = row_addr + min(lab, last) * (NW * 4u);
        if constexpr (NW == 1) {
            uint32_t v;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
            acc[d][0] += v;
        } else if constexpr (NW == 2) {
            uint32_t v0, v1;
            asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v0), "=r"(v1) : "r"(a));
            acc[d][0] += v0; acc[d][1] += v1;
        } else {
#pragma unroll
            for (int w = 0; w < NW; w += 4) {
                uint32_t v0, v1, v2, v3;
                asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3) : "r"(a + 4u * w));
                acc[d][w] += v0; acc[d][w + 1] += v1; acc[d][w + 2] += v2; acc[d][w + 3] += v3;
            }
        }
    };
    const uint8_t* const col = indicator + (size_t)row0 * row_stride + i0;
    if (i0 + 8 <= n_data) {
        // four labels per row from two aligned 32-bit loads (the second one is the neighbour's first:
        // an L1 hit) funnel-shifted by the row's misalignment -- rows may start at any byte
        int j = 0;
        for (; j + PI_UNROLL <= rows; j += PI_UNROLL) {
            uint32_t quad[PI_UNROLL];
#pragma unroll
            for (int u = 0; u < PI_UNROLL; ++u) {
                const uint8_t* const p = col + (size_t)(j + u) * row_stride;
                const unsigned shift = (unsigned)(reinterpret_cast<uintptr_t>(p) & 3u);
                const uint32_t* const q = reinterpret_cast<const uint32_t*>(p - shift);
                quad[u] = __funnelshift_r(q[0], q[1], 8u * shift);
            }
            const uint32_t base = table_addr + (uint32_t)j * row_bytes;
#pragma unroll
            for (int u = 0; u < PI_UNROLL; ++u)
#pragma unroll
                for (int d = 0; d < 4; ++d) add(d, base + (uint32_t)u * row_bytes, __byte_perm(quad[u], 0u, 0x4440u + d));
        }
        for (; j < rows; ++j) {
            const uint8_t* const p = col + (size_t)j * row_stride;
#pragma unroll
            for (int d = 0; d < 4; ++d) add(d, table_addr + (uint32_t)j * row_bytes, p[d]);
        }
    } else {                                                // the last threads of a row: byte loads, bounds checked
        for (int j = 0; j < rows; ++j) {
            const uint8_t* const p = col + (size_t)j * row_stride;
#pragma unroll
            for (int d = 0; d < 4; ++d)
                if (i0 + d < n_data) add(d, table_addr + (uint32_t)j * row_bytes, p[d]);
        }
    }
#pragma unroll
    for (int d = 0; d < 4; ++d) {
        if (i0 + d >= n_data) break;
        int32_t* const out = counts + (size_t)(i0 + d) * n_clusters;
#pragma unroll
        for (int w = 0; w < NW; ++w)
#pragma unroll
            for (int byte = 0; byte < 4; ++byte) {
                const int c = 4 * w + byte;
                const uint32_t v = (acc[d][w] >> (8 * byte)) & 0xffu;
                if (c < n_clusters && v) atomicAdd(&out[c], (int32_t)v);   // chunks of rows add up in global memory
            }
    }
}

typedef void (*pi_kernel_fn)(const uint8_t*, int64_t, int, int, const int8_t*, int, int, int, int32_t*);

}  // namespace brta

extern "C" int brta_pindicator_counts(const uint8_t* indicator, int64_t row_stride, int32_t n_rows, int32_t n_data,
                                      const int8_t* cluster_of, int32_t ncomp, int32_t n_clusters,
                                      int32_t* counts, void* stream)
{
    if (!indicator || !cluster_of || !counts) return brta::fail(BRTA_E_NULL, "brta_pindicator_counts: NULL pointer");
    if (n_rows < 0 || n_data < 0 || row_stride < n_data)
        return brta::fail(BRTA_E_RANGE, "brta_pindicator_counts: n_rows, n_data >= 0 and row_stride >= n_data required");
    if (ncomp < 1 || ncomp > 255 || n_clusters < 1 || n_clusters > BRTA_PINDICATOR_MAX_CLUSTERS)
        return brta::fail(BRTA_E_NCOMP, "brta_pindicator_counts: 1 <= ncomp <= 255, 1 <= n_clusters <= 32");
    if (n_rows == 0 || n_data == 0) return 0;
    const int nw = n_clusters <= 4 ? 1 : n_clusters <= 8 ? 2 : n_clusters <= 16 ? 4 : 8;
    brta::pi_kernel_fn fn = nw == 1 ? brta::pindicator_kernel<1> : nw == 2 ? brta::pindicator_kernel<2>
                          : nw == 4 ? brta::pindicator_kernel<4> : brta::pindicator_kernel<8>;
    // rows per chunk: what fits the table, at most 255 (byte counters), a multiple of the unroll if possible
    int rows_per_chunk = brta::PI_SMEM_TABLE / ((ncomp + 1) * nw * 4);
    rows_per_chunk = rows_per_chunk > 255 ? 255 : rows_per_chunk;
    if (rows_per_chunk >= brta::PI_UNROLL) rows_per_chunk -= rows_per_chunk % brta::PI_UNROLL;
    if (rows_per_chunk < 1) return brta::fail(BRTA_E_RANGE, "brta_pindicator_counts: table does not fit shared memory");
    // few data: split the rows further so that the grid still covers the SMs a few times
    const long long col_blocks = (n_data + 4 * brta::PI_THREADS - 1) / (4 * brta::PI_THREADS);
    while (rows_per_chunk >= 64 && col_blocks * ((n_rows + rows_per_chunk - 1) / rows_per_chunk) < 1184)
        rows_per_chunk = (rows_per_chunk / 2 + brta::PI_UNROLL - 1) / brta::PI_UNROLL * brta::PI_UNROLL;
    const size_t smem = (size_t)rows_per_chunk * (ncomp + 1) * nw * 4;
    const dim3 grid((unsigned)((n_data + 4 * brta::PI_THREADS - 1) / (4 * brta::PI_THREADS)),
                    (unsigned)((n_rows + rows_per_chunk - 1) / rows_per_chunk));
    if (grid.y > 65535u) return brta::fail(BRTA_E_RANGE, "brta_pindicator_counts: too many rows");
    fn<<<grid, brta::PI_THREADS, smem, (cudaStream_t)stream>>>(indicator, row_stride, n_rows, n_data, cluster_of, ncomp,
                                                               n_clusters, rows_per_chunk, counts);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return brta::cuda_fail(e, "pindicator_kernel");
    return 0;
}
