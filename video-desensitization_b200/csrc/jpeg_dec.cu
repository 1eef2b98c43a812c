// jpeg_dec.cu — N1: baseline JPEG DECODE on the device, entropy stage included.
//
// Replaces the near end of the reference frame loop
//   img = cv2.imread(path); cv2.cvtColor(img, cv2.COLOR_BGR2RGB)            (reference combine_detect.py:167-173)
// The file's bytes are copied to the GPU as they are; pixels come out IDENTICAL to cv2.imread's (libjpeg-turbo defaults:
// jpeg_idct_islow, fancy up-sampling, the integer YCbCr->RGB tables) — checked against oracle/jpeg_oracle.py, which is
// pinned on cv2.imdecode.
//
// Stages (all files of a batch at once, no host round trip):
//   unstuff   0xFF00 -> 0xFF: parallel compaction of the entropy-coded segment (count per 1 KB chunk, then scatter)
//   sync      PARALLEL HUFFMAN DECODING by self-synchronisation.  The stream is cut into subsequences of `sub_bytes`;
//             one thread per subsequence decodes from a GUESSED state (bit = its first bit, block 0 of the MCU,
//             coefficient 0).  A Huffman decoder started at a wrong position falls into step with the true one after a
//             few code words; with the JPEG state (block index inside the MCU, zig-zag index) that takes a few hundred
//             bytes.  Every round, thread i re-decodes its subsequence from the state in which thread i-1 LEFT the
//             previous one, and skips the work if that entry state has not changed; subsequence 0 starts from the true
//             state, so after round r the first r+1 exit states are exact, and in practice all of them are after 2-4
//             rounds.  A round in which no exit state changed proves the fixed point (checked; a file that has not
//             converged after kRounds is reported and decoded on the CPU).
//   scan      blocks completed per subsequence -> first block index of every subsequence
//   write     each thread decodes its subsequence once more, now storing coefficients (DC as differences)
//   dc        per component: prefix sum of the DC differences in scan order
//   idct      de-quantisation + jpeg_idct_islow (13-bit constants, column pass then row pass), range limit
//   colour    h2v2 / h2v1 fancy (triangle) up-sampling + YCbCr -> RGB, written as HWC uint8 into the frame
#include "common.cuh"

namespace {

__constant__ uint8_t kZig[80] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                 41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63,
                                 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63};

constexpr int kChunk = 1024;          // unstuff: raw bytes per CTA
constexpr int kRounds = 32;           // synchronisation rounds (the work of a round is skipped where nothing changed):
                                      // exit state i is exact once a wrong-start decoder has run (rounds + 1) x sub_bytes,
                                      // i.e. 16 KB at the default subsequence size of 512 bytes (photographic frames need 1-4 KB)
constexpr int kSubThreads = 128;      // decoder threads per CTA
constexpr int kPad = 64;              // zero bytes behind the unstuffed stream (the bit reader looks 8 bytes ahead)

struct Geo {                          // what every file of the batch shares
  int W, H, ncomp;
  int h[3], v[3], bx[3], by[3];
  int mcux, mcuy, nb;
  uint8_t blk_comp[12], blk_bx[12], blk_by[12];
  int total_blocks;
};

struct DecWs {
  uint8_t* U;                          // [B][ucap] unstuffed entropy data
  uint32_t* ulen;                      // [B]
  uint32_t* uchunk;                    // [B][chunks]
  unsigned long long* exit_state;      // [B][nsub]  state in which subsequence i is left
  unsigned long long* last_entry;      // [B][nsub]  entry state of the decode that produced exit_state[i]
  uint32_t* nblk;                      // [B][nsub]  blocks completed inside subsequence i
  uint32_t* blk_base;                  // [B][nsub]  exclusive scan of nblk
  uint32_t* changed;                   // [B][kRounds]
  int16_t* coef[3];                    // [B][blocks * 64] natural order
  uint8_t* plane[3];                   // [B][by*8][bx*8]
  size_t ucap, coef_elems[3], plane_bytes[3];
  int chunks, nsub, sub_bits;
};

__host__ __device__ inline size_t al256(size_t v) { return (v + 255) & ~static_cast<size_t>(255); }

__host__ Geo make_geo(const vd_jpeg_dev_hdr_t& g) {
  Geo G;
  memset(&G, 0, sizeof(G));
  G.W = g.width; G.H = g.height; G.ncomp = g.ncomp;
  for (int c = 0; c < 3; ++c) { G.h[c] = g.h[c]; G.v[c] = g.v[c]; G.bx[c] = g.blocks_x[c]; G.by[c] = g.blocks_y[c]; }
  G.mcux = g.mcux; G.mcuy = g.mcuy; G.nb = g.blocks_per_mcu;
  memcpy(G.blk_comp, g.blk_comp, 12); memcpy(G.blk_bx, g.blk_bx, 12); memcpy(G.blk_by, g.blk_by, 12);
  G.total_blocks = g.mcux * g.mcuy * g.blocks_per_mcu;
  return G;
}

__host__ size_t carve(void* base, int B, const Geo& G, size_t max_scan, int sub_bytes, DecWs* out) {
  DecWs w;
  memset(&w, 0, sizeof(w));
  uint8_t* p = static_cast<uint8_t*>(base);
  auto take = [&](size_t bytes) { uint8_t* r = p; p += al256(bytes); return r; };
  w.ucap = al256(max_scan + kPad);
  w.chunks = static_cast<int>((max_scan + kChunk - 1) / kChunk);
  w.nsub = static_cast<int>((max_scan + sub_bytes - 1) / sub_bytes) + 1;
  w.sub_bits = sub_bytes * 8;
  w.U = take(static_cast<size_t>(B) * w.ucap);
  w.ulen = reinterpret_cast<uint32_t*>(take(static_cast<size_t>(B) * 4));
  w.uchunk = reinterpret_cast<uint32_t*>(take(static_cast<size_t>(B) * w.chunks * 4));
  w.exit_state = reinterpret_cast<unsigned long long*>(take(static_cast<size_t>(B) * w.nsub * 8));
  w.last_entry = reinterpret_cast<unsigned long long*>(take(static_cast<size_t>(B) * w.nsub * 8));
  w.nblk = reinterpret_cast<uint32_t*>(take(static_cast<size_t>(B) * w.nsub * 4));
  w.blk_base = reinterpret_cast<uint32_t*>(take(static_cast<size_t>(B) * w.nsub * 4));
  w.changed = reinterpret_cast<uint32_t*>(take(static_cast<size_t>(B) * kRounds * 4));
  for (int c = 0; c < 3; ++c) {
    w.coef_elems[c] = c < G.ncomp ? static_cast<size_t>(G.bx[c]) * G.by[c] * 64 : 0;
    w.plane_bytes[c] = c < G.ncomp ? static_cast<size_t>(G.bx[c]) * G.by[c] * 64 : 0;
    w.coef[c] = reinterpret_cast<int16_t*>(take(static_cast<size_t>(B) * w.coef_elems[c] * 2));
    w.plane[c] = take(static_cast<size_t>(B) * w.plane_bytes[c]);
  }
  if (out) *out = w;
  return static_cast<size_t>(p - static_cast<uint8_t*>(base));
}

// ---------------------------------------------------------------------------------------------
// unstuff
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool stuffed_zero(const uint8_t* __restrict__ s, uint32_t i) {
  return i > 0 && s[i] == 0x00 && s[i - 1] == 0xFF;
}

__global__ void __launch_bounds__(256)
unstuff_count_kernel(const vd_jpeg_dev_hdr_t* __restrict__ hdrs, const uint8_t* __restrict__ bytes, DecWs ws) {
  const int f = blockIdx.y;
  const uint32_t n = hdrs[f].scan_bytes;
  const uint8_t* s = bytes + hdrs[f].scan_offset;
  const uint32_t c0 = blockIdx.x * kChunk;
  uint32_t cnt = 0;
  for (uint32_t i = c0 + threadIdx.x; i < min(c0 + kChunk, n); i += 256) cnt += stuffed_zero(s, i);
  cnt = __reduce_add_sync(0xFFFFFFFFu, cnt);
  __shared__ uint32_t sm[8];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) ws.uchunk[static_cast<size_t>(f) * ws.chunks + blockIdx.x] = sm[0] + sm[1] + sm[2] + sm[3] + sm[4] + sm[5] + sm[6] + sm[7];
}

__global__ void __launch_bounds__(256)
unstuff_scatter_kernel(const vd_jpeg_dev_hdr_t* __restrict__ hdrs, const uint8_t* __restrict__ bytes, DecWs ws) {
  const int f = blockIdx.y, tid = threadIdx.x;
  const uint32_t n = hdrs[f].scan_bytes;
  const uint8_t* s = bytes + hdrs[f].scan_offset;
  const uint32_t nchunks = (n + kChunk - 1) / kChunk;
  if (blockIdx.x > nchunks) return;
  const uint32_t* uc = ws.uchunk + static_cast<size_t>(f) * ws.chunks;
  uint32_t before = 0, all = 0;
  for (uint32_t i = tid; i < nchunks; i += 256) {
    const uint32_t v = uc[i];
    all += v;
    if (i < blockIdx.x) before += v;
  }
  before = __reduce_add_sync(0xFFFFFFFFu, before);
  all = __reduce_add_sync(0xFFFFFFFFu, all);
  __shared__ uint32_t sb[8], sa[8], sc[8];
  if ((tid & 31) == 0) { sb[tid >> 5] = before; sa[tid >> 5] = all; }
  __syncthreads();
  before = sb[0] + sb[1] + sb[2] + sb[3] + sb[4] + sb[5] + sb[6] + sb[7];
  all = sa[0] + sa[1] + sa[2] + sa[3] + sa[4] + sa[5] + sa[6] + sa[7];
  uint8_t* U = ws.U + static_cast<size_t>(f) * ws.ucap;
  if (blockIdx.x == nchunks) {                 // one extra CTA: length + zero padding behind the data
    const uint32_t ulen = n - all;
    if (tid == 0) ws.ulen[f] = ulen;
    if (tid < kPad) U[ulen + tid] = 0;
    return;
  }
  const uint32_t c0 = blockIdx.x * kChunk;
  uint32_t rm[4], nrm = 0;
  uint8_t b[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint32_t i = c0 + 4 * tid + k;
    rm[k] = 2;                                 // 2 = beyond the data
    if (i < n) {
      b[k] = s[i];
      rm[k] = stuffed_zero(s, i) ? 1u : 0u;
      nrm += rm[k];
    }
  }
  uint32_t incl = nrm;
  const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) sc[warp] = incl;
  __syncthreads();
  uint32_t wbase = 0;
  for (int w = 0; w < warp; ++w) wbase += sc[w];
  uint32_t o = c0 + 4 * tid - before - (wbase + incl - nrm);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (rm[k] == 2) break;
    if (rm[k] == 0) U[o++] = b[k];
  }
}

// ---------------------------------------------------------------------------------------------
// entropy decoding
// ---------------------------------------------------------------------------------------------
struct HuffS {                 // layout of vd_jpeg_huff_dev_t
  uint16_t fast[512];
  int maxcode[18];
  int valoff[18];
  uint8_t symbols[256];
};
static_assert(sizeof(HuffS) == sizeof(vd_jpeg_huff_dev_t), "HuffS mirrors vd_jpeg_huff_dev_t");

__device__ __forceinline__ uint32_t peek32(const uint8_t* __restrict__ U, uint32_t p) {
  const uint32_t* w = reinterpret_cast<const uint32_t*>(U) + (p >> 5);
  const uint32_t w0 = __byte_perm(w[0], 0, 0x0123), w1 = __byte_perm(w[1], 0, 0x0123);     // big-endian words
  return __funnelshift_l(w1, w0, p & 31);
}

// returns the symbol (or -1: no code matches — a desynchronised decoder; the caller then steps one bit) and its length
__device__ __forceinline__ int huff_symbol(const HuffS& t, uint32_t w, int& len) {
  const uint32_t e = t.fast[w >> 23];
  if (e) {
    len = static_cast<int>(e >> 8);
    return static_cast<int>(e & 0xFF);
  }
  len = 9;
  int code = static_cast<int>(w >> 23);
  while (len < 17 && code > t.maxcode[len]) {
    ++len;
    code = static_cast<int>(w >> (32 - len));
  }
  if (len > 16) {
    len = 1;
    return -1;
  }
  return t.symbols[(t.valoff[len] + code) & 255];
}

__device__ __forceinline__ unsigned long long pack_state(uint32_t p, int b, int k) {
  return static_cast<unsigned long long>(p) | (static_cast<unsigned long long>(b) << 32) | (static_cast<unsigned long long>(k) << 40);
}

// Decodes from state (p, b, k) until p >= end_bit.  WRITE: coefficients of block g (global scan-order index, advanced with
// every completed block) are stored.  Returns the number of blocks completed.
template <bool WRITE>
__device__ __forceinline__ uint32_t decode_range(const uint8_t* __restrict__ U, const HuffS* __restrict__ dc,
                                                 const HuffS* __restrict__ ac, const uint8_t* __restrict__ comp_td,
                                                 const uint8_t* __restrict__ comp_ta, const Geo& G, uint32_t& p, int& b, int& k,
                                                 uint32_t end_bit, uint32_t g, int f, const DecWs& ws, uint32_t* __restrict__ bad) {
  uint32_t done = 0;
  int16_t* dst = nullptr;
  auto locate = [&]() {
    dst = nullptr;
    if (WRITE && g < static_cast<uint32_t>(G.total_blocks)) {
      const int m = static_cast<int>(g) / G.nb, j = static_cast<int>(g) - m * G.nb;
      if (j != b) atomicOr(bad, 4u);                                   // block index and decoder state disagree
      const int c = G.blk_comp[j], my = m / G.mcux, mx = m - my * G.mcux;
      const size_t blk = static_cast<size_t>(my * G.v[c] + G.blk_by[j]) * G.bx[c] + mx * G.h[c] + G.blk_bx[j];
      dst = ws.coef[c] + static_cast<size_t>(f) * ws.coef_elems[c] + blk * 64;
    }
  };
  locate();
  // Bit reader: a 64-bit window that always holds >= 32 valid bits starting at bit p, refilled one big-endian word at a
  // time; the NEXT word is loaded as soon as the previous one has been merged, so its latency hides behind the ~5
  // symbols a word carries.  (The first version read two words from memory for every symbol: ncu showed the decoder
  // threads — ~6 warps per SM, nothing to hide latency with — stalled on those loads 5x as long as they issued.)
  const uint32_t* U32 = reinterpret_cast<const uint32_t*>(U);
  uint32_t nextw = (p >> 5) + 2;
  unsigned long long buf = (static_cast<unsigned long long>(__byte_perm(U32[p >> 5], 0, 0x0123)) << 32) |
                           __byte_perm(U32[(p >> 5) + 1], 0, 0x0123);
  buf <<= (p & 31);
  int valid = 64 - static_cast<int>(p & 31);
  uint32_t ahead = __byte_perm(U32[nextw], 0, 0x0123);
  auto consume = [&](int n) {
    buf <<= n;
    valid -= n;
    p += static_cast<uint32_t>(n);
    if (valid <= 32) {
      buf |= static_cast<unsigned long long>(ahead) << (32 - valid);
      valid += 32;
      ++nextw;
      ahead = p < end_bit ? __byte_perm(U32[nextw], 0, 0x0123) : 0u;
    }
  };
  // One code path for DC and AC symbols (selects instead of branches): the threads of a warp decode unrelated streams.
  // Only the long-code path of huff_symbol, block ends and stores still diverge.
  while (p < end_bit) {
    const uint32_t w = static_cast<uint32_t>(buf >> 32);
    const int c = G.blk_comp[b];
    const bool isdc = k == 0;
    const HuffS& tbl = isdc ? dc[comp_td[c]] : ac[comp_ta[c]];
    int len;
    const int sym = huff_symbol(tbl, w, len);
    if (sym < 0 || (isdc && sym > 15)) { consume(1); continue; }     // no code matches (a desynchronised decoder): step one bit
    const int run = isdc ? 0 : (sym >> 4);
    const int size = isdc ? sym : (sym & 15);
    const bool eob = !isdc && size == 0 && run != 15;
    const bool zrl = !isdc && size == 0 && run == 15;
    const int kc = isdc ? 0 : k + run;                                 // coefficient this symbol carries (if size > 0)
    if (WRITE && dst && size > 0 && kc < 64) {
      const uint32_t v = (w << len) >> (32 - size);
      const int m = static_cast<int>(v >> (size - 1)) - 1;            // T.81 EXTEND
      dst[kZig[kc]] = static_cast<int16_t>(static_cast<int>(v) + (m & (1 - (1 << size))));
    }
    consume(len + size);
    k = eob ? 64 : (zrl ? k + 16 : kc + 1);
    if (k >= 64) {
      k = 0;
      b = b + 1 == G.nb ? 0 : b + 1;
      ++done;
      ++g;
      if (WRITE) locate();
    }
  }
  return done;
}

// one synchronisation round (see the header comment); `first` = every thread starts from its guess
__global__ void __launch_bounds__(kSubThreads)
huff_sync_kernel(const vd_jpeg_dev_hdr_t* __restrict__ hdrs, const __grid_constant__ Geo G, DecWs ws, int round) {
  // The Huffman tables are read straight from the file's record in global memory (5.7 KB, L1-resident): the kernel
  // uses NO shared memory, so its CTAs can co-reside with the detector's 227 KB convolution CTAs and decode the next
  // batch in their shadow instead of queueing between them.
  const int f = blockIdx.y;
  const HuffS* s_dc = reinterpret_cast<const HuffS*>(hdrs[f].dc);
  const HuffS* s_ac = reinterpret_cast<const HuffS*>(hdrs[f].ac);
  const uint8_t* s_t = hdrs[f].td;                 // td[3] is followed by ta[3]
  const uint32_t total_bits = ws.ulen[f] * 8u;
  const int i = blockIdx.x * kSubThreads + threadIdx.x;
  const uint32_t start_bit = static_cast<uint32_t>(i) * static_cast<uint32_t>(ws.sub_bits);
  if (i >= ws.nsub || start_bit >= total_bits) return;
  const size_t idx = static_cast<size_t>(f) * ws.nsub + i;
  unsigned long long entry;
  if (i == 0) entry = pack_state(0, 0, 0);
  else if (round == 0) entry = pack_state(start_bit, 0, 0);
  else entry = ws.exit_state[idx - 1];
  if (round > 0 && entry == ws.last_entry[idx]) return;                     // same entry state: same exit state
  uint32_t p = static_cast<uint32_t>(entry & 0xFFFFFFFFull);
  int b = static_cast<int>((entry >> 32) & 0xFF), k = static_cast<int>((entry >> 40) & 0xFF);
  const uint32_t end_bit = min(start_bit + static_cast<uint32_t>(ws.sub_bits), total_bits);
  const uint32_t done = decode_range<false>(ws.U + static_cast<size_t>(f) * ws.ucap, s_dc, s_ac, s_t, s_t + 3, G, p, b, k,
                                            end_bit, 0u, f, ws, nullptr);
  const unsigned long long ex = pack_state(p, b, k);
  if (round == 0 || ex != ws.exit_state[idx]) atomicAdd(ws.changed + static_cast<size_t>(f) * kRounds + round, 1u);
  ws.exit_state[idx] = ex;
  ws.last_entry[idx] = entry;
  ws.nblk[idx] = done;
}

// exclusive scan of the block counts of one file; flags a file whose decode did not converge or whose block count is off
__global__ void __launch_bounds__(1024)
huff_scan_kernel(const __grid_constant__ Geo G, DecWs ws, int32_t* __restrict__ result) {
  __shared__ uint32_t s_w[32];
  __shared__ uint32_t s_carry;
  const int f = blockIdx.x, tid = threadIdx.x;
  const uint32_t total_bits = ws.ulen[f] * 8u;
  const int nsub = static_cast<int>((total_bits + ws.sub_bits - 1) / ws.sub_bits);
  if (tid == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < nsub; base += 1024) {
    const int i = base + tid;
    const uint32_t v = i < nsub ? ws.nblk[static_cast<size_t>(f) * ws.nsub + i] : 0u;
    uint32_t incl = v;
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t n = __shfl_up_sync(0xFFFFFFFFu, incl, o);
      if (lane >= o) incl += n;
    }
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    uint32_t wbase = s_carry;
    for (int w = 0; w < warp; ++w) wbase += s_w[w];
    if (i < nsub) ws.blk_base[static_cast<size_t>(f) * ws.nsub + i] = wbase + incl - v;
    __syncthreads();
    if (tid == 1023) s_carry = wbase + incl;
    __syncthreads();
  }
  if (tid == 0) {
    // the padding bits behind the last block may decode as one more (bogus) block: tolerated, its writes are dropped
    const uint32_t tot = s_carry;
    const bool count_ok = tot == static_cast<uint32_t>(G.total_blocks) || tot == static_cast<uint32_t>(G.total_blocks) + 1u;
    const bool converged = ws.changed[static_cast<size_t>(f) * kRounds + (kRounds - 1)] == 0;
    int first_quiet = kRounds;                                  // first round in which no exit state changed
    for (int r = kRounds - 1; r >= 1; --r)
      if (ws.changed[static_cast<size_t>(f) * kRounds + r] == 0) first_quiet = r;
    // bits 0..7: 1 = not converged, 2 = block count off (4 = desync, write pass); bits 8..15: rounds it took
    result[f] = (converged ? 0 : 1) | (count_ok ? 0 : 2) | (first_quiet << 8);
  }
}

__global__ void __launch_bounds__(kSubThreads)
huff_write_kernel(const vd_jpeg_dev_hdr_t* __restrict__ hdrs, const __grid_constant__ Geo G, DecWs ws,
                  int32_t* __restrict__ result) {
  const int f = blockIdx.y;
  const HuffS* s_dc = reinterpret_cast<const HuffS*>(hdrs[f].dc);
  const HuffS* s_ac = reinterpret_cast<const HuffS*>(hdrs[f].ac);
  const uint8_t* s_t = hdrs[f].td;
  if ((result[f] & 0xFF) != 0) return;
  const uint32_t total_bits = ws.ulen[f] * 8u;
  const int i = blockIdx.x * kSubThreads + threadIdx.x;
  const uint32_t start_bit = static_cast<uint32_t>(i) * static_cast<uint32_t>(ws.sub_bits);
  if (i >= ws.nsub || start_bit >= total_bits) return;
  const size_t idx = static_cast<size_t>(f) * ws.nsub + i;
  const unsigned long long entry = i == 0 ? pack_state(0, 0, 0) : ws.exit_state[idx - 1];
  uint32_t p = static_cast<uint32_t>(entry & 0xFFFFFFFFull);
  int b = static_cast<int>((entry >> 32) & 0xFF), k = static_cast<int>((entry >> 40) & 0xFF);
  const uint32_t end_bit = min(start_bit + static_cast<uint32_t>(ws.sub_bits), total_bits);
  uint32_t bad = 0;
  decode_range<true>(ws.U + static_cast<size_t>(f) * ws.ucap, s_dc, s_ac, s_t, s_t + 3, G, p, b, k, end_bit,
                     ws.blk_base[idx], f, ws, reinterpret_cast<uint32_t*>(result + f));
  (void)bad;
}

// DC differences -> DC values: prefix sum per component over its blocks in SCAN order
__global__ void __launch_bounds__(1024)
dc_prefix_kernel(const __grid_constant__ Geo G, DecWs ws) {
  __shared__ int s_w[32];
  const int c = blockIdx.x, f = blockIdx.y, tid = threadIdx.x;
  if (c >= G.ncomp) return;
  const int per_mcu = G.h[c] * G.v[c];
  const int n = G.mcux * G.mcuy * per_mcu;
  const int chunk = (n + 1023) / 1024;
  int16_t* coef = ws.coef[c] + static_cast<size_t>(f) * ws.coef_elems[c];
  auto raster = [&](int s) {
    const int m = s / per_mcu, j = s - m * per_mcu;
    const int my = m / G.mcux, mx = m - my * G.mcux;
    const int byy = j / G.h[c], bxx = j - byy * G.h[c];
    return (static_cast<size_t>(my * G.v[c] + byy) * G.bx[c] + mx * G.h[c] + bxx) * 64;
  };
  const int s0 = tid * chunk, s1 = min(n, s0 + chunk);
  int sum = 0;
  for (int s = s0; s < s1; ++s) sum += coef[raster(s)];
  int incl = sum;
  const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) s_w[warp] = incl;
  __syncthreads();
  int run = incl - sum;
  for (int w = 0; w < warp; ++w) run += s_w[w];
  for (int s = s0; s < s1; ++s) {
    const size_t o = raster(s);
    run += coef[o];
    coef[o] = static_cast<int16_t>(run);
  }
}

// ---------------------------------------------------------------------------------------------
// jidctint.c jpeg_idct_islow
// ---------------------------------------------------------------------------------------------
#define JF_0_298 2446
#define JF_0_390 3196
#define JF_0_541 4433
#define JF_0_765 6270
#define JF_0_899 7373
#define JF_1_175 9633
#define JF_1_501 12299
#define JF_1_847 15137
#define JF_1_961 16069
#define JF_2_053 16819
#define JF_2_562 20995
#define JF_3_072 25172

template <int SHIFT>
__device__ __forceinline__ void idct8(const int v[8], int o[8]) {
  int z2 = v[2], z3 = v[6];
  int z1 = (z2 + z3) * JF_0_541;
  const int tmp2 = z1 - z3 * JF_1_847, tmp3 = z1 + z2 * JF_0_765;
  const int tmp0 = (v[0] + v[4]) << 13, tmp1 = (v[0] - v[4]) << 13;
  const int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
  int t0 = v[7], t1 = v[5], t2 = v[3], t3 = v[1];
  z1 = t0 + t3;
  z2 = t1 + t2;
  z3 = t0 + t2;
  int z4 = t1 + t3;
  const int z5 = (z3 + z4) * JF_1_175;
  t0 *= JF_0_298; t1 *= JF_2_053; t2 *= JF_3_072; t3 *= JF_1_501;
  z1 *= -JF_0_899; z2 *= -JF_2_562;
  z3 = z3 * (-JF_1_961) + z5;
  z4 = z4 * (-JF_0_390) + z5;
  t0 += z1 + z3; t1 += z2 + z4; t2 += z2 + z3; t3 += z1 + z4;
  constexpr int R = 1 << (SHIFT - 1);
  o[0] = (tmp10 + t3 + R) >> SHIFT; o[7] = (tmp10 - t3 + R) >> SHIFT;
  o[1] = (tmp11 + t2 + R) >> SHIFT; o[6] = (tmp11 - t2 + R) >> SHIFT;
  o[2] = (tmp12 + t1 + R) >> SHIFT; o[5] = (tmp12 - t1 + R) >> SHIFT;
  o[3] = (tmp13 + t0 + R) >> SHIFT; o[4] = (tmp13 - t0 + R) >> SHIFT;
}

constexpr int kIdctBlocks = 32;       // 8x8 blocks per CTA, eight threads each

__global__ void __launch_bounds__(kIdctBlocks * 8)
idct_kernel(const vd_jpeg_dev_hdr_t* __restrict__ hdrs, const __grid_constant__ Geo G, DecWs ws, int c) {
  __shared__ int s_ws[kIdctBlocks][64];
  __shared__ uint16_t s_q[64];
  const int f = blockIdx.y, tid = threadIdx.x;
  if (tid < 64) s_q[tid] = hdrs[f].quant[c][tid];
  __syncthreads();
  const int nblocks = G.bx[c] * G.by[c];
  const int lb = tid >> 3, t = tid & 7;
  const int blk = blockIdx.x * kIdctBlocks + lb;
  const bool on = blk < nblocks;
  if (on) {                                   // pass 1: column t of the de-quantised block
    const int16_t* src = ws.coef[c] + static_cast<size_t>(f) * ws.coef_elems[c] + static_cast<size_t>(blk) * 64;
    int v[8], o[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) v[r] = static_cast<int>(src[r * 8 + t]) * static_cast<int>(s_q[r * 8 + t]);
    idct8<13 - 2>(v, o);
#pragma unroll
    for (int r = 0; r < 8; ++r) s_ws[lb][r * 8 + t] = o[r];
  }
  __syncthreads();
  if (on) {                                   // pass 2: row t, then the range limit table (centred at 128, 10-bit index)
    int v[8], o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = s_ws[lb][t * 8 + i];
    idct8<13 + 2 + 3>(v, o);
    uint32_t px[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int idx = o[i] & 0x3FF;
      px[i] = static_cast<uint32_t>(idx < 128 ? idx + 128 : (idx < 512 ? 255 : (idx < 896 ? 0 : idx - 896)));
    }
    const int brow = blk / G.bx[c], bcol = blk - brow * G.bx[c];
    uint8_t* dst = ws.plane[c] + static_cast<size_t>(f) * ws.plane_bytes[c] +
                   (static_cast<size_t>(brow) * 8 + t) * (G.bx[c] * 8) + bcol * 8;
    uint2 pk;
    pk.x = px[0] | (px[1] << 8) | (px[2] << 16) | (px[3] << 24);
    pk.y = px[4] | (px[5] << 8) | (px[6] << 16) | (px[7] << 24);
    *reinterpret_cast<uint2*>(dst) = pk;
  }
}

// ---------------------------------------------------------------------------------------------
// jdsample.c fancy up-sampling + jdcolor.c ycc_rgb_convert
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int clamp255(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }

// chroma sample of output pixel (y, x) of a plane with dw x dh real samples and `pitch` bytes per row
__device__ __forceinline__ int upsample_at(const uint8_t* __restrict__ P, int pitch, int dw, int dh, int hs, int vs, int y, int x) {
  if (hs == 1 && vs == 1) return P[static_cast<size_t>(y) * pitch + x];
  if (hs == 2 && vs == 1) {                               // h2v1_fancy_upsample
    const uint8_t* row = P + static_cast<size_t>(y) * pitch;
    const int cc = x >> 1;
    if (dw <= 2) return row[cc];
    const int p = row[cc];
    if (x & 1) return cc == dw - 1 ? p : (3 * p + row[cc + 1] + 2) >> 2;
    return cc == 0 ? p : (3 * p + row[cc - 1] + 1) >> 2;
  }
  // h2v2_fancy_upsample: rows above / below clamped to the component's real extent
  const int r = y >> 1, cc = x >> 1;
  if (dw <= 2) return P[static_cast<size_t>(r) * pitch + cc];
  const int ro = (y & 1) ? min(r + 1, dh - 1) : max(r - 1, 0);
  const uint8_t* a = P + static_cast<size_t>(r) * pitch;
  const uint8_t* o = P + static_cast<size_t>(ro) * pitch;
  const int cs = 3 * a[cc] + o[cc];
  if (x & 1) {
    if (cc == dw - 1) return (4 * cs + 7) >> 4;
    return (3 * cs + (3 * a[cc + 1] + o[cc + 1]) + 7) >> 4;
  }
  if (cc == 0) return (4 * cs + 8) >> 4;
  return (3 * cs + (3 * a[cc - 1] + o[cc - 1]) + 8) >> 4;
}

__global__ void __launch_bounds__(256)
colour_kernel(const __grid_constant__ Geo G, DecWs ws, const vd_frame_t* __restrict__ frames_out) {
  const int f = blockIdx.z, y = blockIdx.y;
  const vd_frame_t fr = frames_out[f];
  const int W = G.W, H = G.H;
  const int py = G.bx[0] * 8;
  const uint8_t* Y = ws.plane[0] + static_cast<size_t>(f) * ws.plane_bytes[0] + static_cast<size_t>(y) * py;
  uint8_t* out = const_cast<uint8_t*>(fr.ptr) + static_cast<size_t>(y) * fr.pitch;
  if (G.ncomp == 1) {
    for (int x = blockIdx.x * 256 + threadIdx.x; x < W; x += gridDim.x * 256) {
      const uint8_t g = Y[x];
      out[3 * x] = g; out[3 * x + 1] = g; out[3 * x + 2] = g;
    }
    return;
  }
  const int hs = G.h[0], vs = G.v[0];                     // chroma is 1x1: pixels per chroma sample
  const int dw = (W + hs - 1) / hs, dh = (H + vs - 1) / vs;
  const int pc = G.bx[1] * 8;
  const uint8_t* Cb = ws.plane[1] + static_cast<size_t>(f) * ws.plane_bytes[1];
  const uint8_t* Cr = ws.plane[2] + static_cast<size_t>(f) * ws.plane_bytes[2];
  for (int x = blockIdx.x * 256 + threadIdx.x; x < W; x += gridDim.x * 256) {
    const int yy = Y[x];
    const int cb = upsample_at(Cb, pc, dw, dh, hs, vs, y, x) - 128;
    const int cr = upsample_at(Cr, pc, dw, dh, hs, vs, y, x) - 128;
    // jdcolor.c build_ycc_rgb_table: FIX(x) = (int)(x * 65536 + 0.5), ONE_HALF = 32768
    const int r = yy + ((91881 * cr + 32768) >> 16);
    const int g = yy + ((-22554 * cb + 32768 - 46802 * cr) >> 16);
    const int b = yy + ((116130 * cb + 32768) >> 16);
    out[3 * x] = static_cast<uint8_t>(clamp255(r));
    out[3 * x + 1] = static_cast<uint8_t>(clamp255(g));
    out[3 * x + 2] = static_cast<uint8_t>(clamp255(b));
  }
}

int check_geometry(const vd_jpeg_dev_hdr_t* g) {
  VD_REQUIRE(g, "vd_jpeg: null geometry");
  VD_REQUIRE(g->width > 0 && g->height > 0 && (g->ncomp == 1 || g->ncomp == 3), "vd_jpeg: bad geometry");
  VD_REQUIRE(g->blocks_per_mcu >= 1 && g->blocks_per_mcu <= 6 && g->mcux > 0 && g->mcuy > 0, "vd_jpeg: bad MCU layout");
  return VD_OK;
}

int reconstruct(const vd_jpeg_dev_hdr_t* hdrs, const Geo& G, const DecWs& ws, int B, const vd_frame_t* frames_out,
                cudaStream_t st) {
  for (int c = 0; c < G.ncomp; ++c) {
    idct_kernel<<<dim3(vd_cdiv(G.bx[c] * G.by[c], kIdctBlocks), B), kIdctBlocks * 8, 0, st>>>(hdrs, G, ws, c);
    VD_LAUNCH_CHECK();
  }
  colour_kernel<<<dim3(vd_cdiv(G.W, 1024), G.H, B), 256, 0, st>>>(G, ws, frames_out);
  VD_LAUNCH_CHECK();
  return VD_OK;
}

}  // namespace

extern "C" size_t vd_jpeg_decode_workspace_bytes(int B, const vd_jpeg_dev_hdr_t* geometry, size_t max_scan_bytes) {
  if (B <= 0 || !geometry || check_geometry(geometry) != VD_OK) return 0;
  const Geo G = make_geo(*geometry);
  return carve(nullptr, B, G, max_scan_bytes, 64, nullptr) + 4096;       // sized for the smallest subsequence allowed
}

extern "C" int vd_jpeg_decode_batch(const vd_jpeg_dev_hdr_t* hdrs, const vd_jpeg_dev_hdr_t* geometry_host,
                                    const uint8_t* bytes, int B, size_t max_scan_bytes, const vd_frame_t* frames_out,
                                    int32_t* result, int sub_bytes, void* workspace, size_t workspace_bytes,
                                    vd_stream_t stream) {
  VD_REQUIRE(hdrs && bytes && frames_out && result && workspace, "vd_jpeg_decode_batch: null argument");
  VD_REQUIRE(B > 0 && B <= 65535 && max_scan_bytes > 0 && max_scan_bytes < (1ull << 28), "vd_jpeg_decode_batch: bad extent");
  if (check_geometry(geometry_host) != VD_OK) return VD_ERR_INVALID;
  if (sub_bytes <= 0) sub_bytes = 512;
  VD_REQUIRE(sub_bytes >= 64 && sub_bytes <= 65536 && sub_bytes % 4 == 0, "vd_jpeg_decode_batch: sub_bytes must be a multiple of 4 in [64, 65536]");
  VD_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "vd_jpeg_decode_batch: workspace must be 256-B aligned");
  const Geo G = make_geo(*geometry_host);
  VD_REQUIRE(G.H <= 65535, "vd_jpeg_decode_batch: frame too tall");
  DecWs ws;
  const size_t need = carve(workspace, B, G, max_scan_bytes, sub_bytes, &ws);
  VD_REQUIRE(need <= workspace_bytes, "vd_jpeg_decode_batch: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  VD_CHECK_CUDA(cudaMemsetAsync(ws.changed, 0, static_cast<size_t>(B) * kRounds * 4, st));
  for (int c = 0; c < G.ncomp; ++c)
    VD_CHECK_CUDA(cudaMemsetAsync(ws.coef[c], 0, static_cast<size_t>(B) * ws.coef_elems[c] * 2, st));
  unstuff_count_kernel<<<dim3(ws.chunks, B), 256, 0, st>>>(hdrs, bytes, ws);
  VD_LAUNCH_CHECK();
  unstuff_scatter_kernel<<<dim3(ws.chunks + 1, B), 256, 0, st>>>(hdrs, bytes, ws);
  VD_LAUNCH_CHECK();
  const dim3 sgrid(vd_cdiv(ws.nsub, kSubThreads), B);
  for (int r = 0; r < kRounds; ++r) {
    huff_sync_kernel<<<sgrid, kSubThreads, 0, st>>>(hdrs, G, ws, r);
    VD_LAUNCH_CHECK();
  }
  huff_scan_kernel<<<B, 1024, 0, st>>>(G, ws, result);
  VD_LAUNCH_CHECK();
  huff_write_kernel<<<sgrid, kSubThreads, 0, st>>>(hdrs, G, ws, result);
  VD_LAUNCH_CHECK();
  dc_prefix_kernel<<<dim3(G.ncomp, B), 1024, 0, st>>>(G, ws);
  VD_LAUNCH_CHECK();
  return reconstruct(hdrs, G, ws, B, frames_out, st);
}

extern "C" int vd_jpeg_reconstruct_batch(const vd_jpeg_dev_hdr_t* hdrs, const vd_jpeg_dev_hdr_t* geometry_host, int B,
                                         const int16_t* coef0, const int16_t* coef1, const int16_t* coef2,
                                         const vd_frame_t* frames_out, void* workspace, size_t workspace_bytes,
                                         vd_stream_t stream) {
  VD_REQUIRE(hdrs && coef0 && frames_out && workspace, "vd_jpeg_reconstruct_batch: null argument");
  if (check_geometry(geometry_host) != VD_OK) return VD_ERR_INVALID;
  const Geo G = make_geo(*geometry_host);
  VD_REQUIRE(G.ncomp == 1 || (coef1 && coef2), "vd_jpeg_reconstruct_batch: three coefficient planes needed");
  VD_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "vd_jpeg_reconstruct_batch: workspace must be 256-B aligned");
  DecWs ws;
  const size_t need = carve(workspace, B, G, 4096, 512, &ws);
  VD_REQUIRE(need <= workspace_bytes, "vd_jpeg_reconstruct_batch: workspace too small");
  ws.coef[0] = const_cast<int16_t*>(coef0);
  ws.coef[1] = const_cast<int16_t*>(coef1);
  ws.coef[2] = const_cast<int16_t*>(coef2);
  return reconstruct(hdrs, G, ws, B, frames_out, static_cast<cudaStream_t>(stream));
}
