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

// ------------------------------- encoder -----------------------------------
// Thread per chunk: codewords (len << 56 | code) appended MSB first into a 64-bit
// accumulator, full words stored as they complete.  CB_SHARED: the codebook is
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
  unsigned fill = 0; // bits used in acc
  auto put = [&](unsigned s) {
    const u64 cw = CB_SHARED ? s_cb[s] : __ldg(codebook + s);
    const unsigned len = (unsigned)(cw >> 56);
    if (len) {
      const u64 code = cw & 0x00ffffffffffffffull;
      const unsigned room = 64 - fill;
      if (len < room) {
        acc |= code << (room - len);
        fill += len;
      } else {
        const unsigned rem = len - room;
        *dst++ = acc | (code >> rem);
        acc = rem ? code << (64 - rem) : 0ull;
        fill = rem;
      }
    }
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

} // namespace serial
