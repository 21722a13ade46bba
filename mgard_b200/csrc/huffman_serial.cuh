// Chunk-serial Huffman kernels (sm_100a): ONE THREAD per chunk.
//
// The stream format (reference Lossless/ParallelHuffman/Huffman.hpp:163-239) cuts
// the symbol array into chunks of huff_block_size symbols; every chunk starts on a
// 64-bit word boundary, so chunks are independent bit strings.  The reference
// deflates / decodes one chunk per thread too (Deflate.hpp:46-77,
// Decode.hpp:66-116), with 8-byte symbols and codewords.  At B200 sizes a
// sub-domain has tens of thousands of chunks (2049^2 x 257 fp32: 52 700), i.e. a
// dozen resident warps per SM with NO cross-thread dependency: no self-
// synchronisation passes, no block scans, no shared-memory staging of the stream.
// What makes it run at memory speed here:
//   * 16-bit symbols, code lengths + symbols through a 4096-entry LUT in shared
//     memory (canonical walk only for codes longer than 12 bits);
//   * the bit stream of a chunk lives in a four-word register queue that is
//     refilled three words ahead (the DRAM latency of the next sector is covered
//     by the decoding of the words in hand);
//   * 256-bit global accesses (LDG/STG.E.ENL2.256, sm_100+): a thread reads 16
//     symbols / writes 8 dequantized fp32 values as ONE full 32-byte sector, so
//     the uncoalesced-by-construction access pattern still moves whole sectors.
// The block-per-chunk kernels of huffman.cu stay for inputs with few chunks.
#pragma once
#include <type_traits>

namespace serial {

constexpr int DS_T = 512;          // most threads (= chunks) per block, decoder (chosen per launch)
constexpr int ES_T = 128;          // encoder
constexpr int LUT_BITS = 16;       // code lengths resolved by one table lookup

// Decoder tables in global memory, built once per block of the stream and copied to
// shared memory by every thread block:
//   len8[x]   x = next 16 stream bits: length of the codeword they start with
//             (0: longer than 16 bits)                                  65536 B
//   key16[k]  symbols in canonical order (decodebook keys)              dict x 2 B
//   t32[l]    first[l] left aligned in 32 bits (0xffffffff: no code of length l)
//   b32[l]    entry[l] - first[l] (mod 2^32): key index = b32[l] + code
// layout: t32[36] | b32[36] | len8 | key16
__host__ __device__ inline size_t tab_bytes(int dict) {
  return 72 * 4 + (size_t)(1 << LUT_BITS) + (((size_t)dict * 2 + 15) & ~(size_t)15);
}

__global__ void __launch_bounds__(1024)
build_lut_kernel(const u64 *__restrict__ decodebook, int dict, unsigned char *__restrict__ g) {
  __shared__ u64 s_first[64], s_entry[64];
  const int tid = threadIdx.x;
  if (tid < 128)
    (tid < 64 ? s_first : s_entry)[tid & 63] = decodebook[tid];
  __syncthreads();
  unsigned *t32 = reinterpret_cast<unsigned *>(g), *b32 = t32 + 36;
  unsigned char *len8 = g + 72 * 4;
  uint16_t *key16 = reinterpret_cast<uint16_t *>(len8 + (1 << LUT_BITS));
  if (tid < 36) {
    const bool valid = tid >= 1 && tid <= 32 && s_first[tid] != ~0ull && (s_first[tid] >> tid) == 0;
    t32[tid] = valid ? (unsigned)(s_first[tid] << (32 - tid)) : 0xffffffffu;
    b32[tid] = valid ? (unsigned)(s_entry[tid] - s_first[tid]) : 0u;
  }
  int lmin = 1;
  while (lmin < 63 && s_first[lmin] == ~0ull)
    lmin++;
  for (int x = tid; x < (1 << LUT_BITS); x += 1024) {
    unsigned char e = 0;
    for (int l = lmin; l <= LUT_BITS; l++) {
      const u64 v = (u64)x >> (LUT_BITS - l);
      if (v >= s_first[l]) {
        e = (unsigned char)l;
        break;
      }
    }
    len8[x] = e;
  }
  for (int k = tid; k < dict; k += 1024)
    key16[k] = (uint16_t)decodebook[128 + k];
}

template <typename OUT> struct Store8;
template <> struct Store8<float> {
  static __device__ __forceinline__ void run(float *p, const float (&v)[8]) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]),
                 "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
                 : "memory");
  }
};
template <> struct Store8<double> {
  static __device__ __forceinline__ void run(double *p, const double (&v)[8]) {
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3])
                 : "memory");
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p + 4), "d"(v[4]), "d"(v[5]), "d"(v[6]), "d"(v[7])
                 : "memory");
  }
};
template <> struct Store8<uint16_t> {
  static __device__ __forceinline__ void run(uint16_t *p, const uint16_t (&v)[8]) {
    uint4 q;
    q.x = v[0] | ((unsigned)v[1] << 16);
    q.y = v[2] | ((unsigned)v[3] << 16);
    q.z = v[4] | ((unsigned)v[5] << 16);
    q.w = v[6] | ((unsigned)v[7] << 16);
    *reinterpret_cast<uint4 *>(p) = q;
  }
};

template <typename OUT> __device__ __forceinline__ OUT sym_value(unsigned sym, OUT scale, int half) {
  // (quantizer * volume) * (T)(quantized - dict / 2): dequantize_linear_kernel
  return scale * (OUT)((long long)sym - half);
}
template <> __device__ __forceinline__ uint16_t sym_value<uint16_t>(unsigned sym, uint16_t, int) {
  return (uint16_t)sym;
}

__device__ __forceinline__ unsigned lds_u8(unsigned addr) {
  unsigned v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ unsigned lds_u16(unsigned addr) {
  unsigned v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ unsigned lds_u32(unsigned addr) {
  unsigned v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}

// VEC: chunk starts are 32-byte aligned in `out` (8 values per store)
template <typename OUT, bool VEC>
__global__ void __launch_bounds__(DS_T)
decode_serial_kernel(const u64 *__restrict__ ddata, u64 total_words, const u64 *__restrict__ bits,
                     const u64 *__restrict__ woff, u64 nchunk, int chunk, u64 n,
                     const u64 *__restrict__ decodebook, int dict, const unsigned char *__restrict__ gtab,
                     OUT *__restrict__ out, OUT scale) {
  extern __shared__ __align__(16) unsigned char s_tab[];
  {
    const uint4 *g4 = reinterpret_cast<const uint4 *>(gtab);
    uint4 *s4 = reinterpret_cast<uint4 *>(s_tab);
    const int n16 = (int)(tab_bytes(dict) / 16);
    for (int i = threadIdx.x; i < n16; i += blockDim.x)
      s4[i] = g4[i];
  }
  __syncthreads();
  const u64 c = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nchunk)
    return;
  // shared addresses as plain 32-bit registers (no generic-pointer arithmetic in the loop)
  const unsigned a_t32 = (unsigned)__cvta_generic_to_shared(s_tab), a_b32 = a_t32 + 36 * 4;
  const unsigned a_len = a_t32 + 72 * 4, a_key = a_len + (1 << LUT_BITS);
  const int half = dict / 2;
  const unsigned nsym = (unsigned)min((u64)chunk, n - c * (u64)chunk);
  OUT *dst = out + c * (u64)chunk;
  const u64 B64 = bits[c], w0 = woff[c];
  const u64 nw = (B64 - 1) / 64 + 1;
  // the per-chunk fields come from the stream: a chunk that does not lie inside the
  // bit stream decodes to zeros (as in decode_kernel)
  if (B64 == 0 || B64 > (u64)chunk * 64 || w0 > total_words || nw > total_words - w0) {
    for (unsigned i = 0; i < nsym; i++)
      dst[i] = (OUT)0;
    return;
  }
  const u64 *src = ddata + w0;
  auto ldw = [&](u64 i) -> u64 { return i < nw ? __ldg(src + i) : 0ull; };
  // the sectors of this chunk's bit stream are requested 256 bytes ahead of the reader
  auto prefetch = [&](u64 i) {
    if (i < nw)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(src + i));
  };
  u64 cur = ldw(0), nxt = ldw(1), n2 = ldw(2), n3 = ldw(3);
#pragma unroll
  for (int k = 1; k <= 8; k++)
    prefetch(4 * k);
  u64 wi = 0;
  unsigned pos = 0; // bit position inside cur
  const u64 *first = decodebook, *entry = decodebook + 64, *keys = decodebook + 128;

  auto step = [&]() -> unsigned {
    // 32 stream bits starting at bit `pos` of cur:nxt
    const unsigned A = (unsigned)(cur >> 32), Bw = (unsigned)cur, C = (unsigned)(nxt >> 32);
    const bool lowhalf = pos < 32;
    const unsigned hi = __funnelshift_l(lowhalf ? Bw : C, lowhalf ? A : Bw, pos);
    unsigned l = lds_u8(a_len + (hi >> (32 - LUT_BITS)));
    unsigned sym;
    if (l == 0) {
      // longer than 16 bits: first length whose left-aligned first code is <= the window
      l = LUT_BITS + 1;
      while (l <= 32 && hi < lds_u32(a_t32 + l * 4))
        l++;
      if (l > 32) {
        // canonical walk on a 64-bit window (codes of 33 .. 63 bits)
        const u64 win = pos ? ((cur << pos) | (nxt >> (64 - pos))) : cur;
        int ll = 33;
        u64 v = win >> (64 - ll);
        while (v < __ldg(first + ll) && ll < 63) {
          ll++;
          v = win >> (64 - ll);
        }
        const u64 ki = __ldg(entry + ll) + v - __ldg(first + ll);
        sym = ki < (u64)dict ? (unsigned)(__ldg(keys + ki) & 0xffffu) : 0u;
        l = (unsigned)ll;
      } else {
        const unsigned ki = lds_u32(a_b32 + l * 4) + (hi >> (32 - l));
        sym = ki < (unsigned)dict ? lds_u16(a_key + ki * 2) : 0u;
      }
    } else {
      const unsigned ki = lds_u32(a_b32 + l * 4) + (hi >> (32 - l));
      sym = ki < (unsigned)dict ? lds_u16(a_key + ki * 2) : 0u;
    }
    pos += l;
    if (pos >= 64) {
      pos -= 64;
      cur = nxt;
      nxt = n2;
      n2 = n3;
      wi++;
      n3 = ldw(wi + 3);
      if ((wi & 3) == 0)
        prefetch(wi + 36);
    }
    return sym;
  };

  unsigned i = 0;
  if (VEC) {
    for (; i + 8 <= nsym; i += 8) {
      OUT v[8];
#pragma unroll
      for (int k = 0; k < 8; k++)
        v[k] = sym_value<OUT>(step(), scale, half);
      Store8<OUT>::run(dst + i, v);
    }
  }
  for (; i < nsym; i++)
    dst[i] = sym_value<OUT>(step(), scale, half);
}

// ----------------------- decoder, second formulation ------------------------
// Thread per chunk again, written so that the lanes of a warp never part ways
// although each is at a different bit position of a different chunk — a warp pays
// for a branch whenever ANY of its lanes takes it, and with 32 independent decoders
// in a warp "rare" events (a new word of the bit stream, a codeword that misses the
// table) happen at almost every symbol — and with as few instructions per symbol as
// it takes: with one warp per 32 chunks there are only ~3 warps per scheduler, each
// a chain of dependent instructions, so the instruction count IS the run time.
//   * The reader holds 64 stream bits (w0:w1), the next 32 (w2) and a bit offset
//     o < 32; the 32-bit window at o is ONE funnel shift.  After a codeword: o += l,
//     and if o >= 32 the words move up and w2 is reloaded — predicated, no branch.
//   * Codewords of up to RL_BITS bits: one lookup by the first RL_BITS bits of the
//     window: lut[x] = {dequantized value, length} (fp32 output) or symbol << 8 | length.
//   * Codewords of up to RL_BITS + RL2_BITS bits: a miss of the first table carries
//     the offset of a second-level table for its prefix (codes longer than RL_BITS
//     bits are the numerically smallest ones: their prefixes are 0, 1, 2, ...), looked
//     up with the following RL2_BITS bits by a PREDICATED load.
//   * Longer codewords (a few in 10^4 symbols): canonical walk on the window, out of
//     line; beyond 32 bits (never seen): from global memory.
//   * The bit stream of a thread's chunk reaches it through a private 128-byte ring
//     in shared memory that the thread keeps full itself with 16-byte cp.async copies,
//     RING - 1 pieces ahead of its reader (pieces and words XOR-swizzled by the lane,
//     so the lanes of a warp hit different banks).
// Consecutive chunks go to different thread blocks and warps: chunks of one part of
// the coefficient array are alike, and the ones that hold the coarse levels (long
// codewords throughout) would otherwise share a warp.
// Requires 32-byte aligned chunk starts in `out` (8 values per store).
constexpr int RL_BITS = 12;
constexpr int RL2_BITS = 8;        // second-level tables: up to RL2_BITS more bits
constexpr int RL2_SHALLOW = 4;     // ... or RL2_SHALLOW where no codeword of the prefix needs more
constexpr int RING = 8;            // 16-byte pieces per thread
constexpr size_t RL2_MAX_BYTES = 160 * 1024; // second-level tables at most

template <typename OUT> struct RingLut { typedef unsigned entry; };
template <> struct RingLut<float> { typedef uint2 entry; };

// Tables (E = table entry): lut[1 << RL_BITS] | t32[36] | b32[36] | lstart[36] | pad[4] | key16[dict] | sub[sub_bytes]
//   lut     first level; a miss (length 0) carries where its prefix continues: byte offset of
//           a second-level table in `sub` and 32 - d, d = number of index bits of that table
//           (RL2_BITS, RL2_SHALLOW if every codeword of the prefix has at most RL_BITS +
//           RL2_SHALLOW bits, 0 for the one-entry table 0 that misses: no room left)
//   sub     second level, same entries, indexed by the d bits that follow the prefix
//   t32[l]  first code of length l left aligned in 32 bits (0xffffffff: none), b32[l] =
//           entry[l] - first[l], lstart[z] = shortest length a 32-bit window with z leading
//           zeros can have: where the canonical walk starts; key16: the symbols in canonical order
__host__ __device__ inline size_t ring_key_bytes(int dict) { return ((size_t)dict * 2 + 127) & ~(size_t)127; }
template <typename OUT> __host__ __device__ inline size_t ring_tab_bytes(int dict, size_t sub_bytes) {
  return (sizeof(typename RingLut<OUT>::entry) << RL_BITS) + 112 * 4 + ring_key_bytes(dict) + sub_bytes;
}
template <typename OUT> __host__ __device__ inline size_t ring_smem_bytes(int dict, size_t sub_bytes, int slots) {
  return ring_tab_bytes<OUT>(dict, sub_bytes) + (size_t)RING * 16 * slots;
}

template <typename OUT>
__global__ void __launch_bounds__(1024)
build_ring_lut_kernel(const u64 *__restrict__ decodebook, int dict, unsigned sub_bytes, unsigned char *__restrict__ g,
                      OUT scale) {
  typedef typename RingLut<OUT>::entry E;
  __shared__ u64 s_first[64], s_entry[64];
  __shared__ unsigned char s_depth[1 << RL_BITS]; // index bits of the prefix's second-level table (0: a hit)
  __shared__ unsigned s_off[1 << RL_BITS];        // its byte offset in `sub`
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < 128)
    (tid < 64 ? s_first : s_entry)[tid & 63] = decodebook[tid];
  __syncthreads();
  E *lut = reinterpret_cast<E *>(g);
  unsigned *t32 = reinterpret_cast<unsigned *>(lut + (1 << RL_BITS)), *b32 = t32 + 36, *lstart = b32 + 36;
  uint16_t *key16 = reinterpret_cast<uint16_t *>(lstart + 40);
  unsigned char *sub = reinterpret_cast<unsigned char *>(key16) + ring_key_bytes(dict);
  auto t32_of = [&](int l) -> unsigned {
    const bool valid = l >= 1 && l <= 32 && s_first[l] != ~0ull && (s_first[l] >> l) == 0;
    return valid ? (unsigned)(s_first[l] << (32 - l)) : 0xffffffffu;
  };
  if (tid < 36) {
    t32[tid] = t32_of(tid);
    b32[tid] = t32_of(tid) != 0xffffffffu ? (unsigned)(s_entry[tid] - s_first[tid]) : 0u;
    // windows with z = tid leading zeros are at most hi_max: no length whose first code lies above it
    const unsigned hi_max = tid >= 32 ? 0u : (0xffffffffu >> tid);
    int l = RL_BITS + 1;
    while (l <= 32 && t32_of(l) > hi_max)
      l++;
    lstart[tid] = (unsigned)l;
  }
  for (int k = tid; k < dict; k += 1024)
    key16[k] = (uint16_t)decodebook[128 + k];
  int lmin = 1;
  while (lmin < 63 && s_first[lmin] == ~0ull)
    lmin++;
  const int half = dict / 2;
  // symbol << 8 | length of the codeword at the top of the `width`-bit value x (0: longer)
  auto lookup = [&](unsigned x, int width) -> unsigned {
    for (int l = lmin; l <= width; l++) {
      const u64 v = (u64)x >> (width - l);
      if (v >= s_first[l]) {
        const u64 ki = s_entry[l] + v - s_first[l];
        const unsigned sym = ki < (u64)dict ? (unsigned)(decodebook[128 + ki] & 0xffffu) : 0u;
        return (sym << 8) | (unsigned)l;
      }
    }
    return 0u;
  };
  // table entry of a codeword (e != 0), or of a miss that continues in the table of 1 << d
  // entries at byte offset `off`
  auto entry_of = [&](unsigned e, unsigned off, unsigned d) -> E {
    if constexpr (sizeof(E) == 8) {
      const float val = scale * (float)((int)(e >> 8) - half);
      return (e & 0xffu) ? make_uint2(__float_as_uint(val), e & 0xffu) : make_uint2(off | ((32 - d) << 24), 0u);
    } else {
      return (e & 0xffu) ? e : (((off >> 2) | ((32 - d) << 18)) << 8);
    }
  };
  // which prefixes miss, and how deep their tables have to be (a warp per prefix)
  for (int x = tid; x < (1 << RL_BITS); x += 1024)
    s_depth[x] = (lookup((unsigned)x, RL_BITS) & 0xffu) ? 0 : 0xff;
  __syncthreads();
  for (int x = warp; x < (1 << RL_BITS); x += 32) {
    if (s_depth[x] != 0xff)
      continue;
    bool deep = false;
    for (int j = lane; j < (1 << RL2_BITS); j += 32) {
      const unsigned len = lookup(((unsigned)x << RL2_BITS) | (unsigned)j, RL_BITS + RL2_BITS) & 0xffu;
      deep = deep || len == 0 || len > RL_BITS + RL2_SHALLOW;
    }
    deep = __any_sync(0xffffffffu, deep);
    __syncwarp();
    if (lane == 0)
      s_depth[x] = deep ? RL2_BITS : RL2_SHALLOW;
  }
  __syncthreads();
  // offsets (block-wide exclusive sum of the table sizes, four prefixes per thread): table 0
  // is the single entry that misses; a prefix whose table does not fit gets it
  {
    __shared__ unsigned s_warp[32];
    unsigned size[4], mine = 0;
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const unsigned d = s_depth[4 * tid + q];
      size[q] = d ? (unsigned)sizeof(E) << d : 0u;
      mine += size[q];
    }
    unsigned incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o)
        incl += t;
    }
    if (lane == 31)
      s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      unsigned w = s_warp[lane], wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, wi, o);
        if (lane >= o)
          wi += t;
      }
      s_warp[lane] = wi - w;
    }
    __syncthreads();
    unsigned run = 16 + s_warp[warp] + incl - mine;
#pragma unroll
    for (int q = 0; q < 4; q++) {
      if (size[q]) {
        if (run + size[q] <= sub_bytes) {
          s_off[4 * tid + q] = run;
        } else {
          s_off[4 * tid + q] = 0;
          s_depth[4 * tid + q] = 0xfe; // marks "table 0"
        }
      }
      run += size[q];
    }
  }
  __syncthreads();
  if (tid < 16 / (int)sizeof(E))
    reinterpret_cast<E *>(sub)[tid] = entry_of(0u, 0u, 0u);
  for (int x = tid; x < (1 << RL_BITS); x += 1024) {
    const unsigned d = s_depth[x];
    lut[x] = d == 0 ? entry_of(lookup((unsigned)x, RL_BITS), 0u, 0u) : d == 0xfe ? entry_of(0u, 0u, 0u) : entry_of(0u, s_off[x], d);
  }
  for (int x = warp; x < (1 << RL_BITS); x += 32) {
    const unsigned d = s_depth[x];
    if (d == 0 || d == 0xfe)
      continue;
    E *tab = reinterpret_cast<E *>(sub + s_off[x]);
    for (int j = lane; j < (1 << d); j += 32) {
      const unsigned e = lookup(((unsigned)x << d) | (unsigned)j, RL_BITS + (int)d);
      tab[j] = entry_of((e & 0xffu) > RL_BITS ? e : 0u, 0u, 0u); // a miss here: longer than RL_BITS + d bits
    }
  }
}

// helpers of the ring decoder's cold paths (inlined: a call in the kernel makes the compiler
// keep what lives across it in local memory, which the hot loop then reads back)
// 32 stream bits starting at half-word h of the chunk whose words are [A, lim), counted from A16
__device__ __forceinline__ unsigned ring_half_global(unsigned h, unsigned long long A16, unsigned long long A,
                                                  unsigned long long lim) {
  const unsigned long long a = A16 + (u64)(h >> 1) * 8 + ((h & 1) ^ 1) * 4;
  return (a >= A && a < lim) ? __ldg(reinterpret_cast<const unsigned *>(a)) : 0u;
}
// codeword of RL_BITS + 1 .. 32 bits at the top of `hi`: symbol | length << 16; ~0: longer
__device__ __forceinline__ unsigned ring_walk(unsigned hi, unsigned a_t32, unsigned a_key, int dict) {
  const unsigned a_b32 = a_t32 + 36 * 4, a_lstart = a_b32 + 36 * 4;
  unsigned l = lds_u32(a_lstart + __clz(hi) * 4);
#pragma unroll 1
  while (l <= 32 && hi < lds_u32(a_t32 + l * 4))
    l++;
  if (l > 32)
    return ~0u;
  const unsigned ki = lds_u32(a_b32 + l * 4) + (hi >> (32 - l));
  const unsigned sym = ki < (unsigned)dict ? lds_u16(a_key + ki * 2) : 0u;
  return sym | (l << 16);
}
// codeword of 33 .. 63 bits at stream bit `bitpos` (from A16): symbol | length << 16
__device__ __forceinline__ unsigned ring_walk_long(u64 bitpos, unsigned long long A16, unsigned long long A,
                                                unsigned long long lim, const u64 *__restrict__ decodebook,
                                                int dict) {
  const unsigned h = (unsigned)(bitpos >> 5), off = (unsigned)(bitpos & 31);
  const u64 x0 = ((u64)ring_half_global(h, A16, A, lim) << 32) | ring_half_global(h + 1, A16, A, lim);
  const u64 win = off ? ((x0 << off) | ((u64)ring_half_global(h + 2, A16, A, lim) >> (32 - off))) : x0;
  const u64 *first = decodebook, *entry = decodebook + 64, *keys = decodebook + 128;
  int ll = 33;
  u64 v = win >> (64 - ll);
  while (v < __ldg(first + ll) && ll < 63) {
    ll++;
    v = win >> (64 - ll);
  }
  const u64 ki = __ldg(entry + ll) + v - __ldg(first + ll);
  const unsigned sym = ki < (u64)dict ? (unsigned)(__ldg(keys + ki) & 0xffffu) : 0u;
  return sym | ((unsigned)ll << 16);
}

constexpr int RING_T = 768;        // most threads per block
constexpr int RING_S = 32;         // the lanes of a warp take every RING_S-th chunk
template <typename OUT>
__global__ void __launch_bounds__(RING_T)
decode_ring_kernel(const u64 *__restrict__ ddata, u64 total_words, const u64 *__restrict__ bits,
                   const u64 *__restrict__ woff, u64 nchunk, int chunk, u64 n,
                   const u64 *__restrict__ decodebook, int dict, unsigned sub_bytes, int lanes,
                   const unsigned char *__restrict__ gtab, OUT *__restrict__ out, OUT scale) {
  typedef typename RingLut<OUT>::entry E;
  constexpr bool VAL_LUT = sizeof(E) == 8; // the tables hold dequantized values
  constexpr int ES = VAL_LUT ? 3 : 2;      // log2 of the entry size
  extern __shared__ __align__(128) unsigned char s_tab[];
  {
    const uint4 *g4 = reinterpret_cast<const uint4 *>(gtab);
    uint4 *s4 = reinterpret_cast<uint4 *>(s_tab);
    const int n16 = (int)(ring_tab_bytes<OUT>(dict, sub_bytes) / 16);
    for (int i = threadIdx.x; i < n16; i += blockDim.x)
      s4[i] = g4[i];
  }
  __syncthreads();
  // Chunk of this thread.  Only the first `lanes` lanes of a warp take one (32: all of them).
  // Chunks of one part of the coefficient array are alike, and the ones that hold the coarse
  // levels (long codewords throughout, a few dozen in a row) should not share a warp: warps
  // work in groups of RING_S, the lanes of a warp take every RING_S-th chunk of the group's
  // lanes * RING_S chunks.  (Spreading further costs more than it gains: the 32 stores of a
  // warp then go to 32 pages that are not in the TLB.)
  const unsigned lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const u64 wg = (u64)blockIdx.x * (blockDim.x >> 5) + wib;
  const u64 c = (wg / RING_S) * ((u64)lanes * RING_S) + (u64)lane * RING_S + wg % RING_S;
  if (lane >= (unsigned)lanes || c >= nchunk)
    return;
  // shared addresses held in registers (opaque to the compiler, which otherwise
  // recomputes the window base in front of every access)
  unsigned a_lut = (unsigned)__cvta_generic_to_shared(s_tab);
  asm volatile("mov.u32 %0, %0;" : "+r"(a_lut));
  const unsigned a_t32 = a_lut + (unsigned)(sizeof(E) << RL_BITS);
  const unsigned a_key = a_t32 + 112 * 4;
  unsigned a_sub = a_key + (unsigned)ring_key_bytes(dict);
  // this thread's ring: 128 bytes; piece q at ((q ^ lane) & 7) * 16, half-word h (32 stream
  // bits, the HIGH half of a 64-bit word first) at ((4 * h) ^ cx) & 124
  unsigned a_ring = a_lut + (unsigned)ring_tab_bytes<OUT>(dict, sub_bytes) + (wib * lanes + lane) * (RING * 16u);
  unsigned cx = 4u ^ ((lane & 7u) << 4);
  asm volatile("mov.u32 %0, %0;" : "+r"(a_sub));
  asm volatile("mov.u32 %0, %0;" : "+r"(a_ring));
  asm volatile("mov.u32 %0, %0;" : "+r"(cx));
  const int half = dict / 2;
  const unsigned nsym = (unsigned)min((u64)chunk, n - c * (u64)chunk);
  OUT *dst = out + c * (u64)chunk;
  const u64 B64 = bits[c], w0_ = woff[c];
  const u64 nw = (B64 - 1) / 64 + 1;
  // the per-chunk fields come from the stream: a chunk that does not lie inside the
  // bit stream decodes to zeros (as in decode_kernel)
  if (B64 == 0 || B64 > (u64)chunk * 64 || w0_ > total_words || nw > total_words - w0_) {
    for (unsigned i = 0; i < nsym; i++)
      dst[i] = (OUT)0;
    return;
  }
  // byte addresses: the chunk's words are [A, lim); pieces and half-words are counted from A16
  const unsigned long long A = (unsigned long long)(ddata + w0_), lim = A + nw * 8, A16 = A & ~15ull;
  unsigned iss = 0; // pieces requested so far: the ring is full after every request phase
  // piece `iss` if it is below `top` (no branch)
  auto request = [&](unsigned top) {
    const unsigned long long g = A16 + (u64)iss * 16;
    const unsigned nbytes = g < lim ? (unsigned)min((unsigned long long)16, lim - g) : 0u;
    const unsigned go = iss < top;
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %3, 0;\n\t@p cp.async.cg.shared.global [%0], [%1], 16, %2;\n\t}" ::"r"(
                     a_ring + (((iss ^ lane) & 7u) << 4)),
                 "l"(nbytes ? g : A16), "r"(nbytes), "r"(go)
                 : "memory");
    iss += go;
  };
  auto half_ring = [&](unsigned hb) -> unsigned { // hb = 4 * half-word index
    unsigned v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a_ring + ((hb ^ cx) & 124u)) : "memory");
    return v;
  };
#pragma unroll
  for (int q = 0; q < RING; q++)
    request(RING);
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  // A round (8 symbols) ends with the ring full (RING pieces from the one the reader's next
  // word is in) and all but the requests of the last two rounds complete.  A round takes
  // at most 8 * 32 bits = 2 pieces, so what is read during the next round was requested
  // three rounds ago or earlier: always there.  (Codewords of more than 32 bits: see `rare`.)
  unsigned w0, w1, w2; // stream bits: the window is at bit o of w0:w1; w2 follows
  unsigned o;          // < 32
  unsigned hb;         // 4 * (half-word that follows w2)
  {
    const unsigned h0 = (unsigned)(A - A16) / 4;
    w0 = half_ring(4 * h0), w1 = half_ring(4 * h0 + 4), w2 = half_ring(4 * h0 + 8);
    hb = 4 * h0 + 12;
    o = 0;
  }
  auto value = [&](unsigned sym) -> OUT {
    if (sizeof(OUT) == 2)
      return (OUT)sym;
    return scale * (OUT)((int)sym - half); // = (T)(long long)(sym - half): |sym - half| < 2^16
  };
  // a codeword longer than RL_BITS + RL2_BITS bits: symbol and length
  auto rare = [&](unsigned win, unsigned &l) -> unsigned {
    unsigned r = ring_walk(win, a_t32, a_key, dict);
    if (r == ~0u) {
      // more than 32 bits: from global memory, and the reader restarts behind it
      const u64 bitpos = (u64)(hb / 4 - 3) * 32 + o;
      r = ring_walk_long(bitpos, A16, A, lim, decodebook, dict);
      const u64 np = bitpos + (r >> 16);
      const unsigned h = (unsigned)(np >> 5);
      w0 = ring_half_global(h, A16, A, lim), w1 = ring_half_global(h + 1, A16, A, lim);
      w2 = ring_half_global(h + 2, A16, A, lim);
      hb = 4 * h + 12;
      o = (unsigned)(np & 31);
      asm volatile("cp.async.wait_group 0;" ::: "memory"); // everything requested is there
      l = 0;
      return r & 0xffffu;
    }
    l = r >> 16;
    return r & 0xffffu;
  };
  // One symbol.  FAST: a codeword that is in neither table leaves the reader where it is
  // (length 0) and is not counted in `ndec`; the caller deals with it.
  unsigned ndec = 0;
  auto step = [&](auto fast) -> OUT {
    constexpr bool FAST = decltype(fast)::value;
    const unsigned win = __funnelshift_l(w1, w0, o);
    unsigned l, e0;
    OUT v;
    // first level, and the second one predicated on its miss: the entry then holds the
    // offset of the prefix's table and 32 - (its index bits); PTX shifts by 32 give 0
    const unsigned w12 = win << RL_BITS;
    unsigned idx;
    if constexpr (VAL_LUT) {
      asm("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(e0), "=r"(l) : "r"(a_lut + ((win >> (32 - RL_BITS)) << 3)));
      asm("shr.u32 %0, %1, %2;" : "=r"(idx) : "r"(w12), "r"(e0 >> 24));
      asm("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, %1, 0;\n\t@p ld.shared.v2.u32 {%0, %1}, [%2];\n\t}"
          : "+r"(e0), "+r"(l)
          : "r"(a_sub + (e0 & 0xffffffu) + (idx << 3)));
      v = __uint_as_float(e0);
    } else {
      e0 = lds_u32(a_lut + ((win >> (32 - RL_BITS)) << 2));
      asm("shr.u32 %0, %1, %2;" : "=r"(idx) : "r"(w12), "r"(e0 >> 26));
      asm("{\n\t.reg .pred p;\n\t.reg .u32 t;\n\tand.b32 t, %0, 255;\n\tsetp.eq.u32 p, t, 0;\n\t@p ld.shared.u32 %0, [%1];\n\t}"
          : "+r"(e0)
          : "r"(a_sub + ((e0 >> 6) & 0xffffcu) + (idx << 2)));
      l = e0 & 0xffu;
      v = value(e0 >> 8);
    }
    if (FAST)
      ndec += l != 0;
    else if (l == 0)
      v = value(rare(win, l));
    // behind the codeword; a new word when the offset leaves w0
    o += l;
    const bool p = o >= 32;
    o = p ? o - 32 : o;
    w0 = p ? w1 : w0;
    w1 = p ? w2 : w1;
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p ld.shared.u32 %0, [%1];\n\t}"
                 : "+r"(w2)
                 : "r"(a_ring + ((hb ^ cx) & 124u)), "r"((unsigned)p)
                 : "memory");
    hb += p ? 4u : 0u;
    return v;
  };

  unsigned i = 0;
  for (; i + 8 <= nsym; i += 8) {
    OUT v[8];
    ndec = 0;
#pragma unroll
    for (int k = 0; k < 8; k++)
      v[k] = step(std::true_type());
    Store8<OUT>::run(dst + i, v);
    if (__builtin_expect(ndec < 8, 0)) {
      // the reader stands at a long codeword: this one and the rest of the round one by one
#pragma unroll 1
      for (unsigned k = ndec; k < 8; k++)
        dst[i + k] = step(std::false_type());
    }
    // the pieces below the one the next word is in have been read: their slots take the next
    // pieces (two at most: a round takes no more)
    const unsigned top = (hb >> 4) + RING;
    request(top);
    request(top);
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 2;" ::: "memory");
  }
  for (; i < nsym; i++)
    dst[i] = step(std::false_type());
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// ------------------------------- encoder -----------------------------------
// Thread per chunk: codewords (len << 56 | code) appended MSB first into a 64-bit
// accumulator, full words stored as they complete (predicated, no branch).  CB_SHARED: the codebook is
// copied to shared memory (dict * 8 bytes); otherwise it is read through L1.
template <bool CB_SHARED, bool VEC>
__global__ void __launch_bounds__(ES_T)
encode_serial_kernel(const uint16_t *__restrict__ sym, u64 n, int chunk, const u64 *__restrict__ codebook,
                     int dict, const u64 *__restrict__ woff, const u64 *__restrict__ scal,
                     u64 *__restrict__ ddata) {
  extern __shared__ u64 s_cb[];
  if (scal[2])
    return; // output too small / outlier overflow: nothing is written
  if (CB_SHARED) {
    for (int i = threadIdx.x; i < dict; i += ES_T)
      s_cb[i] = codebook[i];
    __syncthreads();
  }
  const u64 nchunk = (n - 1) / chunk + 1;
  const u64 c = (u64)blockIdx.x * ES_T + threadIdx.x;
  if (c >= nchunk)
    return;
  const u64 lo = c * (u64)chunk;
  const unsigned cnt = (unsigned)min((u64)chunk, n - lo);
  const uint16_t *src = sym + lo;
  u64 *dst = ddata + woff[c];
  u64 acc = 0;
  unsigned fill = 0; // bits used in acc (< 64)
  // Appends the codeword of s without a branch (the lanes of a warp complete their words at
  // different symbols: a branch would be taken by some lane at almost every one).  The
  // codeword goes to bit `fill` of the 128-bit pair acc : next word; PTX shifts by 64 or
  // more give 0, which selects the part that applies.
  auto put = [&](unsigned s) {
    const u64 cw = CB_SHARED ? s_cb[s] : __ldg(codebook + s);
    const unsigned len = (unsigned)(cw >> 56);
    const u64 code = cw & 0x00ffffffffffffffull;
    const int sh = 64 - (int)fill - (int)len; // >= 0: the codeword ends inside acc
    u64 hi1, hi2, lo;
    asm("shl.b64 %0, %1, %2;" : "=l"(hi1) : "l"(code), "r"((unsigned)sh));
    asm("shr.b64 %0, %1, %2;" : "=l"(hi2) : "l"(code), "r"((unsigned)-sh));
    asm("shl.b64 %0, %1, %2;" : "=l"(lo) : "l"(code), "r"((unsigned)(64 + sh)));
    const u64 out = acc | hi1 | hi2;
    const unsigned nf = fill + len;
    const unsigned full = nf >= 64;
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p st.global.u64 [%0], %1;\n\t}" ::"l"(dst), "l"(out),
                 "r"(full)
                 : "memory");
    dst += full;
    acc = full ? lo : out;
    fill = full ? nf - 64 : nf;
  };
  unsigned i = 0;
  if (VEC && cnt >= 16) {
    // 16 symbols = one 32-byte sector per load; the next sector is already in
    // registers while this one is packed, the ones after it are on their way to L2
    auto load16 = [&](unsigned at, unsigned (&w)[8]) {
      asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
                   : "l"(src + at));
    };
    auto prefetch = [&](unsigned at) {
      if (at < cnt)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(src + at));
    };
    unsigned w[8], wn[8];
    load16(0, w);
#pragma unroll
    for (int k = 1; k <= 8; k++)
      prefetch(16 * k);
    for (; i + 16 <= cnt; i += 16) {
      if (i + 32 <= cnt)
        load16(i + 16, wn);
      prefetch(i + 16 * 9);
#pragma unroll
      for (int k = 0; k < 16; k++)
        put((w[k >> 1] >> (16 * (k & 1))) & 0xffffu);
#pragma unroll
      for (int k = 0; k < 8; k++)
        w[k] = wn[k];
    }
  }
  for (; i < cnt; i++)
    put(src[i]);
  if (fill)
    *dst = acc;
}

// ---------------------- encoder, eight threads per chunk --------------------
// A chunk's bit string is sequential only because every codeword starts where the one
// before it ends; with the bit count of each EIGHTH of a chunk known (a by-product of
// the pass that sizes the chunks) eight threads write one chunk, each the words its
// eighth ends in.  A word that straddles two eighths belongs to the later thread, which
// first re-encodes the few symbols before its start whose codewords reach into that word
// (no atomics, no zero-filled output).  Eight times the threads of encode_serial_kernel:
// the latencies of one chunk's chain are hidden by the other chains of the SM instead of
// setting the run time.
constexpr int ESUB = 8;     // threads per chunk
constexpr int ESUB_T = 256; // threads per block

// bits of every eighth of every chunk (sub[c * ESUB + q]) and of the chunk (bits[c]);
// a block per chunk at a time, warp q sums eighth q.  chunk % (8 * ESUB) == 0.
__global__ void __launch_bounds__(256)
chunk_bits_sub_kernel(const uint16_t *__restrict__ sym, u64 n, int chunk, const u64 *__restrict__ codebook, int dict,
                      u64 *__restrict__ bits, unsigned *__restrict__ sub) {
  extern __shared__ unsigned char s_len[];
  __shared__ unsigned s_part[ESUB];
  for (int i = threadIdx.x; i < dict; i += blockDim.x)
    s_len[i] = (unsigned char)(codebook[i] >> 56);
  __syncthreads();
  const int lane = threadIdx.x & 31, q = threadIdx.x >> 5;
  const u64 nchunk = (n - 1) / chunk + 1;
  const unsigned sublen = (unsigned)chunk / ESUB;
  for (u64 c = blockIdx.x; c < nchunk; c += gridDim.x) {
    const u64 lo = c * (u64)chunk;
    const unsigned cnt = (unsigned)min((u64)chunk, n - lo);
    const unsigned a = min(cnt, q * sublen), b = min(cnt, (q + 1) * sublen);
    const uint16_t *base = sym + lo;
    unsigned total = 0, done = a;
    if ((((uintptr_t)base) & 15) == 0) {
      for (unsigned i = a + 8 * lane; i + 8 <= b; i += 256) {
        const uint4 v = __ldg(reinterpret_cast<const uint4 *>(base + i));
        total += s_len[v.x & 0xffffu] + s_len[v.x >> 16] + s_len[v.y & 0xffffu] + s_len[v.y >> 16] +
                 s_len[v.z & 0xffffu] + s_len[v.z >> 16] + s_len[v.w & 0xffffu] + s_len[v.w >> 16];
      }
      done = a + (b - a) / 8 * 8;
    }
    for (unsigned i = done + lane; i < b; i += 32)
      total += s_len[base[i]];
    for (int o = 16; o > 0; o >>= 1)
      total += __shfl_xor_sync(0xffffffffu, total, o);
    if (lane == 0) {
      s_part[q] = total;
      sub[c * ESUB + q] = total;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      u64 t = 0;
      for (int k = 0; k < ESUB; k++)
        t += s_part[k];
      bits[c] = t;
    }
    __syncthreads();
  }
}

template <bool CB_SHARED>
__global__ void __launch_bounds__(ESUB_T)
encode_sub_kernel(const uint16_t *__restrict__ sym, u64 n, int chunk, const u64 *__restrict__ codebook, int dict,
                  const u64 *__restrict__ woff, const unsigned *__restrict__ sub, const u64 *__restrict__ scal,
                  u64 *__restrict__ ddata) {
  extern __shared__ u64 s_cb[];
  if (scal[2])
    return; // output too small / outlier overflow: nothing is written
  if (CB_SHARED) {
    for (int i = threadIdx.x; i < dict; i += ESUB_T)
      s_cb[i] = codebook[i];
    __syncthreads();
  }
  const u64 nchunk = (n - 1) / chunk + 1;
  const u64 gid = (u64)blockIdx.x * ESUB_T + threadIdx.x;
  const u64 c = gid / ESUB;
  const unsigned q = (unsigned)(gid % ESUB);
  if (c >= nchunk)
    return;
  const u64 lo = c * (u64)chunk;
  const unsigned cnt = (unsigned)min((u64)chunk, n - lo), sublen = (unsigned)chunk / ESUB;
  const unsigned start = min(cnt, q * sublen), end = min(cnt, (q + 1) * sublen);
  // bit range [S, E) of this eighth inside the chunk; T: bits of the chunk
  u64 S = 0, T = 0;
  unsigned mine = 0;
#pragma unroll
  for (unsigned k = 0; k < ESUB; k++) {
    const unsigned b = sub[c * ESUB + k];
    S += k < q ? b : 0u;
    mine = k == q ? b : mine;
    T += b;
  }
  if (mine == 0)
    return; // nothing of this eighth in the stream (beyond the end of the last chunk)
  const u64 E = S + mine;
  const bool last = E == T; // the final, partial word of the chunk is this thread's
  const unsigned w_s = (unsigned)(S >> 6), need = (unsigned)(S & 63);
  const uint16_t *src = sym + lo;
  auto cw_of = [&](unsigned s) -> u64 { return CB_SHARED ? s_cb[s] : __ldg(codebook + s); };
  // the symbols before `start` whose codewords reach into word w_s
  unsigned i0 = start, back = 0;
  while (back < need) {
    i0--;
    back += (unsigned)(cw_of(src[i0]) >> 56);
  }
  const u64 B0 = S - back; // where symbol i0 starts
  u64 *const base = ddata + woff[c];
  unsigned wi = (unsigned)(B0 >> 6); // word being filled
  u64 acc = 0;
  unsigned fill = (unsigned)(B0 & 63); // bits used in acc (< 64); the ones before B0 are not this thread's to write
  // as in encode_serial_kernel; a completed word below w_s (the one symbol i0 starts in) is not stored
  auto put = [&](unsigned s) {
    const u64 cw = cw_of(s);
    const unsigned len = (unsigned)(cw >> 56);
    const u64 code = cw & 0x00ffffffffffffffull;
    const int sh = 64 - (int)fill - (int)len; // >= 0: the codeword ends inside acc
    u64 hi1, hi2, lo2;
    asm("shl.b64 %0, %1, %2;" : "=l"(hi1) : "l"(code), "r"((unsigned)sh));
    asm("shr.b64 %0, %1, %2;" : "=l"(hi2) : "l"(code), "r"((unsigned)-sh));
    asm("shl.b64 %0, %1, %2;" : "=l"(lo2) : "l"(code), "r"((unsigned)(64 + sh)));
    const u64 out = acc | hi1 | hi2;
    const unsigned nf = fill + len;
    const unsigned full = nf >= 64;
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p st.global.u64 [%0], %1;\n\t}" ::"l"(base + wi), "l"(out),
                 "r"(full && wi >= w_s ? 1u : 0u)
                 : "memory");
    wi += full;
    acc = full ? lo2 : out;
    fill = full ? nf - 64 : nf;
  };
  for (unsigned i = i0; i < start; i++)
    put(src[i]);
  unsigned i = start;
  if ((((uintptr_t)(src + start)) & 31) == 0 && end - start >= 16) {
    auto load16 = [&](unsigned at, unsigned (&w)[8]) {
      asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
                   : "l"(src + at));
    };
    unsigned w[8], wn[8];
    load16(i, w);
    for (; i + 16 <= end; i += 16) {
      if (i + 32 <= end)
        load16(i + 16, wn);
#pragma unroll
      for (int k = 0; k < 16; k++)
        put((w[k >> 1] >> (16 * (k & 1))) & 0xffffu);
#pragma unroll
      for (int k = 0; k < 8; k++)
        w[k] = wn[k];
    }
  }
  for (; i < end; i++)
    put(src[i]);
  if (last && fill)
    base[wi] = acc;
}

} // namespace serial
