// Per-datum cluster membership counts from the stored label rows: SURVEY.md 8(f-1).
//
// basicrta/gibbs.py:264-268 walks, for every retained sample row j and every active component k
// of that row, over np.where(indicator[j] == k) and adds 1 to pindicator[i, cluster(j, k)] -- a
// double Python loop over ~1000 rows x active components with an O(N) scan each.  Here it is one
// pass over the label tensor: thread = four data, rows streamed with coalesced 32-bit loads, the
// (row, component) -> increment table of a row chunk in shared memory, byte counters packed in registers.
// Integer work, HBM-bound: n_rows * n_data bytes in, 4 * n_data * n_clusters bytes out.
#include <stdint.h>
#include <stdlib.h>
#include <type_traits>

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

// ---- aligned-class kernel ------------------------------------------------------------------------------------
// The kernel above spends most of its issue slots on the ALU pipe (ncu: ALU 69 %, issue 61 %, 10 instructions per
// label): two 32-bit loads and a funnel shift per row because rows start at any byte, and per label a byte
// extract, a clamp, an address computation, the table load and the add.  This kernel removes most of that:
//   * rows whose start has the same alignment modulo 4 form a *class* (row j has alignment (base + j stride) & 3:
//     periodic in j with period P = 1, 2 or 4); a block walks rows of ONE class, so a thread reads one aligned
//     32-bit word per row and always sees the same four data (shifted by the class offset against the
//     word grid; the few bytes of neighbouring rows it sees at the row ends are masked at the output);
//   * the increment table of a row is a 256-byte aligned block of 256 bytes in shared memory and the four
//     labels of a word are pre-shifted once (label * entry size stays inside its byte), so the address of an
//     entry is ONE byte permute (PRMT: the shifted label replaces the low byte of the row's table address);
//     a label that does not fit (>= 64 / 32 / 16 for 1 / 2 / 4 counter words) sends its word through a slow path.
// Per label: PRMT + LDS + the add.
constexpr int PF_THREADS = 512;
constexpr int PF_UNROLL = 8;                                 // granularity of the row chunks (host side)
// rows in flight per thread: the kernel waits for the row loads (ncu: half of the stall samples on the long
// scoreboard), so one counter word per datum runs 16 deep (0.61 -> 0.635 of the HBM peak); with two or four
// words the registers are better spent elsewhere (8 deep: 0.376 vs 0.371)
__host__ __device__ constexpr int pf_depth(int nw) { return nw == 1 ? 16 : 8; }
constexpr int PF_ROW_BYTES = 256;                            // table bytes per row

template <int NW>
__global__ void __launch_bounds__(PF_THREADS, 3)
pindicator_class_kernel(const uint8_t* __restrict__ indicator, int64_t row_stride, int n_rows, int n_data,
                        const int8_t* __restrict__ cluster_of, int ncomp, int n_clusters, int period, int rows_per_chunk,
                        int n_chunks, int32_t* __restrict__ counts)
{
    extern __shared__ __align__(16) uint8_t pf_raw[];
    constexpr int ENTRY = 4 * NW;                           // bytes per table entry
    constexpr int SHIFT = NW == 1 ? 2 : NW == 2 ? 3 : 4;    // label -> byte offset inside the row's table
    constexpr uint32_t BAD = NW == 1 ? 0xC0C0C0C0u : NW == 2 ? 0xE0E0E0E0u : 0xF0F0F0F0u;   // label * ENTRY >= 256
    const int tid = threadIdx.x;
    const int cls = blockIdx.y / n_chunks, chunk = blockIdx.y - cls * n_chunks;
    const int class_rows = cls < n_rows ? (n_rows - cls + period - 1) / period : 0;      // rows cls, cls + period, ...
    const int m0 = chunk * rows_per_chunk;
    const int rows = min(rows_per_chunk, class_rows - m0);
    if (rows <= 0) return;
    uint32_t table = (uint32_t)__cvta_generic_to_shared(pf_raw);
    table = (table + (PF_ROW_BYTES - 1)) & ~(uint32_t)(PF_ROW_BYTES - 1);
    // zero the table, then fill the entries of the components that are counted
    for (int x = tid; x < rows * (PF_ROW_BYTES / 16); x += PF_THREADS)
        asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" :: "r"(table + 16u * (uint32_t)x), "r"(0u) : "memory");
    __syncthreads();
    const int kbits = 32 - __clz(ncomp - 1 > 0 ? ncomp - 1 : 1);      // components padded to a power of two: no division
    for (int x = tid; x < (rows << kbits); x += PF_THREADS) {
        const int m = x >> kbits, k = x & ((1 << kbits) - 1);
        if (k >= ncomp) continue;
        const int c = (int)cluster_of[(size_t)(cls + (size_t)(m0 + m) * period) * ncomp + k];
        if (c >= 0 && c < n_clusters)
            asm volatile("st.shared.u32 [%0], %1;" :: "r"(table + (uint32_t)m * PF_ROW_BYTES + (uint32_t)k * ENTRY + 4u * (c >> 2)),
                         "r"(1u << (8 * (c & 3))) : "memory");
    }
    __syncthreads();
    // this thread's aligned word of every row of the class, and the data it holds
    const uint8_t* const first = indicator + (size_t)cls * row_stride;                   // row `cls`
    const int a = (int)(reinterpret_cast<uintptr_t>(first) & 3u);                         // its misalignment = the class's
    const long long w = (long long)blockIdx.x * PF_THREADS + tid;
    const long long i0 = 4 * w - a;                                                       // datum of byte 0 of the word
    if (i0 >= n_data) return;
    const int64_t step = row_stride * period;                                             // a multiple of 4 bytes
    const uint8_t* p = first - a + 4 * w + (size_t)m0 * step;
    uint32_t acc[4][NW];
#pragma unroll
    for (int d = 0; d < 4; ++d)
#pragma unroll
        for (int v = 0; v < NW; ++v) acc[d][v] = 0u;
    auto add_entry = [&](int d, uint32_t addr) {
        if constexpr (NW == 1) {
            uint32_t v;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
            acc[d][0] += v;
        } else if constexpr (NW == 2) {
            uint32_t v0, v1;
            asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v0), "=r"(v1) : "r"(addr));
            acc[d][0] += v0; acc[d][1] += v1;
        } else {
            uint32_t v0, v1, v2, v3;
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3) : "r"(addr));
            acc[d][0] += v0; acc[d][1] += v1; acc[d][2] += v2; acc[d][3] += v3;
        }
    };
    // a word whose labels all fit the table: label * ENTRY replaces the low byte of the row's table address
    auto add_word = [&](uint32_t quad, uint32_t row_addr) {
        const uint32_t q = quad << SHIFT;
#pragma unroll
        for (int d = 0; d < 4; ++d) add_entry(d, __byte_perm(q, row_addr, 0x7650u + d));
    };
    // any word: byte by byte, skipping labels beyond the table (never produced by the sampler)
    auto add_word_checked = [&](uint32_t quad, uint32_t row_addr) {
#pragma unroll
        for (int d = 0; d < 4; ++d) {
            const uint32_t lab = (quad >> (8 * d)) & 0xffu;
            if (lab < (uint32_t)ncomp) add_entry(d, row_addr + lab * ENTRY);
        }
    };
    // U rows at a time: U independent loads in flight per thread, one validity test per batch
    auto batch = [&](auto width, int m) {
        constexpr int U = decltype(width)::value;
        uint32_t quad[U];
        uint32_t any = 0u;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            quad[u] = __ldg(reinterpret_cast<const uint32_t*>(p));
            p += step;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) any |= quad[u];
        const uint32_t base = table + (uint32_t)m * PF_ROW_BYTES;
        if (!(any & BAD)) {
#pragma unroll
            for (int u = 0; u < U; ++u) add_word(quad[u], base + (uint32_t)u * PF_ROW_BYTES);
        } else {
#pragma unroll
            for (int u = 0; u < U; ++u) add_word_checked(quad[u], base + (uint32_t)u * PF_ROW_BYTES);
        }
    };
    int m = 0;
    for (; m + pf_depth(NW) <= rows; m += pf_depth(NW)) batch(std::integral_constant<int, pf_depth(NW)>(), m);
    for (; m + 4 <= rows; m += 4) batch(std::integral_constant<int, 4>(), m);
    for (; m < rows; ++m) {
        add_word_checked(__ldg(reinterpret_cast<const uint32_t*>(p)), table + (uint32_t)m * PF_ROW_BYTES);
        p += step;
    }
#pragma unroll
    for (int d = 0; d < 4; ++d) {
        const long long i = i0 + d;
        if (i < 0 || i >= n_data) continue;                 // bytes of the neighbouring rows at the row ends
        int32_t* const out = counts + (size_t)i * n_clusters;
#pragma unroll
        for (int v = 0; v < NW; ++v)
#pragma unroll
            for (int byte = 0; byte < 4; ++byte) {
                const int c = 4 * v + byte;
                const uint32_t n = (acc[d][v] >> (8 * byte)) & 0xffu;
                if (c < n_clusters && n) atomicAdd(&out[c], (int32_t)n);
            }
    }
}

typedef void (*pf_kernel_fn)(const uint8_t*, int64_t, int, int, const int8_t*, int, int, int, int, int, int32_t*);

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
    // the aligned-class kernel whenever every label * entry size fits a 256-byte table row
    const bool use_generic = getenv("BRTA_PINDICATOR_GENERIC") != nullptr;       // developer knob: A/B measurements
    const int pf_period = n_rows > 1 ? ((row_stride & 3) == 0 ? 1 : (row_stride & 1) == 0 ? 2 : 4) : 1;
    // (few rows per alignment class: the table build and the extra output atomics outweigh the cheaper labels --
    // measured 0.35 vs 0.20 ms at 100 rows x 4e6 data)
    if (!use_generic && nw <= 4 && ncomp * nw * 4 <= brta::PF_ROW_BYTES && n_rows / pf_period >= 64) {
        brta::pf_kernel_fn fk = nw == 1 ? brta::pindicator_class_kernel<1> : nw == 2 ? brta::pindicator_class_kernel<2>
                                                                          : brta::pindicator_class_kernel<4>;
        const int period = pf_period;
        const int class_rows = (n_rows + period - 1) / period;
        int rows_per_chunk = class_rows > 248 ? 248 : class_rows;                 // byte counters; a multiple of the unroll
        const long long words = ((long long)n_data + 3 + 3) / 4;
        const long long col_blocks = (words + brta::PF_THREADS - 1) / brta::PF_THREADS;
        while (rows_per_chunk >= 64 && col_blocks * period * ((class_rows + rows_per_chunk - 1) / rows_per_chunk) < 444)
            rows_per_chunk = (rows_per_chunk / 2 + brta::PF_UNROLL - 1) / brta::PF_UNROLL * brta::PF_UNROLL;
        const int n_chunks = (class_rows + rows_per_chunk - 1) / rows_per_chunk;
        const size_t smem = (size_t)rows_per_chunk * brta::PF_ROW_BYTES + brta::PF_ROW_BYTES;
        if ((long long)period * n_chunks > 65535) return brta::fail(BRTA_E_RANGE, "brta_pindicator_counts: too many rows");
        cudaError_t e = cudaFuncSetAttribute(fk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return brta::cuda_fail(e, "cudaFuncSetAttribute(pindicator_class_kernel)");
        fk<<<dim3((unsigned)col_blocks, (unsigned)(period * n_chunks)), brta::PF_THREADS, smem, (cudaStream_t)stream>>>(
            indicator, row_stride, n_rows, n_data, cluster_of, ncomp, n_clusters, period, rows_per_chunk, n_chunks, counts);
        e = cudaGetLastError();
        if (e != cudaSuccess) return brta::cuda_fail(e, "pindicator_class_kernel");
        return 0;
    }
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
