// jpeg_enc.cu — N1: baseline JPEG ENCODE of the processed frames on the device.
//
// Replaces the far end of the reference frame loop
//   cv2.imwrite(output_path, cv2.cvtColor(image_array, COLOR_RGB2BGR))      (reference combine_detect.py:176-180)
// for frames that are already resident in HBM: what crosses PCIe is the ~0.9 MB file of a 1080p frame instead of its
// 6.2 MB of pixels, and no host core spends 14 ms per frame in libjpeg.  The FILE is byte-identical to what cv2.imwrite
// (libjpeg-turbo, quality 95, 4:2:0, standard Huffman tables, no restart markers, JFIF header) produces from the same
// pixels — checked against oracle/jpeg_enc_oracle.py, which is pinned byte for byte against cv2.imencode.
//
// Integer arithmetic restated from libjpeg's published algorithm (ITU T.81 + the IJG fixed-point choices):
//   colour    Y/Cb/Cr = 16-bit fixed point of the BT.601 matrix, chroma rows with the (0.5 - epsilon) rounding fudge
//   sampling  h2v2 box filter with the alternating bias 1, 2, 1, 2, ...; the row is extended by its last pixel, the image
//             by its last DOWN-SAMPLED row; Y blocks beyond the image inside the last MCU are "dummy blocks" (zero AC,
//             DC of the previous block in MCU order)
//   DCT       LL&M 8x8 integer DCT (13-bit constants, 2 extra bits after the row pass), output scaled by 8
//   quant     round-half-away-from-zero division by 8 * q (q from the Annex K tables scaled by the quality)
//   entropy   T.81 Annex F Huffman coding with the Annex K tables, DC differences per component, 0xFF byte stuffing,
//             1-padding of the last byte
//
// Five kernels per batch, all frames at once, no host round trip:
//   dct      RGB -> level-shifted YCbCr -> FDCT -> quantised coefficients in zig-zag order, [frame][MCU][6][64] int16
//   count    bits of every block (+ exclusive scan inside the CTA, CTA totals)
//   emit     every block writes its code words at its bit offset (atomicOr into a zeroed big-endian word stream)
//   ffcount  0xFF bytes per 1 KB chunk of the stream
//   finish   header + byte-stuffed stream + EOI, byte count per frame
#include "common.cuh"

namespace {

constexpr int kZigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                             41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                             30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

__constant__ uint8_t kZigzagDev[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                       41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                       30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

// ITU T.81 Annex K
const uint8_t kStdLumaQ[64] = {16, 11, 10, 16, 24,  40,  51,  61,  12, 12, 14, 19, 26,  58,  60,  55,
                               14, 13, 16, 24, 40,  57,  69,  56,  14, 17, 22, 29, 51,  87,  80,  62,
                               18, 22, 37, 56, 68,  109, 103, 77,  24, 35, 55, 64, 81,  104, 113, 92,
                               49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99};
const uint8_t kStdChromaQ[64] = {17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99,
                                 99, 99, 47, 66, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99,
                                 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99};
const uint8_t kDcLumaBits[16] = {0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0};
const uint8_t kDcChromaBits[16] = {0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0};
const uint8_t kDcVals[12] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11};
const uint8_t kAcLumaBits[16] = {0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7d};
const uint8_t kAcLumaVals[162] = {
    0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07, 0x22, 0x71,
    0x14, 0x32, 0x81, 0x91, 0xa1, 0x08, 0x23, 0x42, 0xb1, 0xc1, 0x15, 0x52, 0xd1, 0xf0, 0x24, 0x33, 0x62, 0x72,
    0x82, 0x09, 0x0a, 0x16, 0x17, 0x18, 0x19, 0x1a, 0x25, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x34, 0x35, 0x36, 0x37,
    0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59,
    0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x83,
    0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3,
    0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3,
    0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe1, 0xe2,
    0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf1, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa};
const uint8_t kAcChromaBits[16] = {0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 0x77};
const uint8_t kAcChromaVals[162] = {
    0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71, 0x13, 0x22,
    0x32, 0x81, 0x08, 0x14, 0x42, 0x91, 0xa1, 0xb1, 0xc1, 0x09, 0x23, 0x33, 0x52, 0xf0, 0x15, 0x62, 0x72, 0xd1,
    0x0a, 0x16, 0x24, 0x34, 0xe1, 0x25, 0xf1, 0x17, 0x18, 0x19, 0x1a, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x35, 0x36,
    0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58,
    0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a,
    0x82, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a,
    0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba,
    0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda,
    0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa};

constexpr int kHeaderMax = 640;
constexpr int kMcusPerCta = 8;            // dct kernel: 8 MCUs = 128 x 16 pixels per CTA
constexpr int kDctThreads = 256;
constexpr int kBlocksPerCta = 256;        // count / emit kernels: one thread per 8x8 block
constexpr int kChunk = 1024;              // ffcount / finish kernels: stream bytes per CTA

// everything a frame size + quality determines, passed by value (constant bank)
struct EncParams {
  int W, H, mcux, mcuy;
  int byY, bxY;                    // real (non-dummy) luma block rows / columns
  int header_len;
  uint8_t header[kHeaderMax];
  unsigned long long magic[2][64];  // ceil(2^40 / (8 q)) per coefficient (natural order): exact quotient for |c| < 2^18
  uint16_t div[2][64];              // 8 q
  // Huffman code tables: (code << 8) | length; DC indexed by category, AC by (run << 4) | size
  uint32_t dc_code[2][12];
  uint32_t ac_code[2][256];
};

struct EncWs {
  int16_t* coef;          // [B][nmcu][6][64] zig-zag order
  uint32_t* blk_off;      // [B][nblk_pad] bit offset of a block inside its CTA
  uint32_t* cta_bits;     // [B][ctas_per_frame]
  uint32_t* words;        // [B][cap_words] big-endian bit stream, zeroed
  uint32_t* ff_count;     // [B][chunks_cap]
  uint32_t* total_bits;   // [B]
  size_t cap_words;
  int nblk, ctas_per_frame, chunks_cap;
};

__host__ __device__ inline size_t al256(size_t v) { return (v + 255) & ~static_cast<size_t>(255); }

__host__ EncWs carve(void* base, int B, int H, int W, size_t total, size_t* zero_from, size_t* zero_bytes) {
  EncWs w;
  const int mcux = vd_cdiv(W, 16), mcuy = vd_cdiv(H, 16);
  w.nblk = mcux * mcuy * 6;
  w.ctas_per_frame = vd_cdiv(w.nblk, kBlocksPerCta);
  uint8_t* p = static_cast<uint8_t*>(base);
  w.coef = reinterpret_cast<int16_t*>(p);
  p += al256(static_cast<size_t>(B) * w.nblk * 64 * 2);
  w.blk_off = reinterpret_cast<uint32_t*>(p);
  p += al256(static_cast<size_t>(B) * w.ctas_per_frame * kBlocksPerCta * 4);
  w.cta_bits = reinterpret_cast<uint32_t*>(p);
  p += al256(static_cast<size_t>(B) * w.ctas_per_frame * 4);
  w.total_bits = reinterpret_cast<uint32_t*>(p);
  p += al256(static_cast<size_t>(B) * 4);
  const size_t used = static_cast<size_t>(p - static_cast<uint8_t*>(base));
  // what is left: the word stream (zeroed every call) and one counter per 1 KB of it
  const size_t left = total > used + 4096 ? total - used - 4096 : 0;
  const size_t per_frame = left / B;
  size_t cap_bytes = (per_frame * kChunk / (kChunk + 4)) & ~static_cast<size_t>(kChunk - 1);
  w.cap_words = cap_bytes / 4;
  w.chunks_cap = static_cast<int>(cap_bytes / kChunk);
  w.words = reinterpret_cast<uint32_t*>(p);
  *zero_from = used;
  *zero_bytes = al256(static_cast<size_t>(B) * cap_bytes);
  p += *zero_bytes;
  w.ff_count = reinterpret_cast<uint32_t*>(p);
  return w;
}

// ---------------------------------------------------------------------------------------------
// 1. colour conversion + down-sampling + FDCT + quantisation
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void rgb_to_ycc(int r, int g, int b, int& y, int& cb, int& cr) {
  // jccolor.c rgb_ycc_convert: FIX(x) = (int)(x * 65536 + 0.5)
  y = (19595 * r + 38470 * g + 7471 * b + 32768) >> 16;
  cb = (-11059 * r - 21709 * g + 32768 * b + (128 << 16) + 32767) >> 16;
  cr = (32768 * r - 27439 * g - 5329 * b + (128 << 16) + 32767) >> 16;
}

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

__device__ __forceinline__ int descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }

// one 1-D pass of jfdctint.c (first = row pass: even outputs scaled up by 4; else the column pass)
template <bool FIRST>
__device__ __forceinline__ void fdct8(const int d[8], int o[8]) {
  const int tmp0 = d[0] + d[7], tmp7 = d[0] - d[7], tmp1 = d[1] + d[6], tmp6 = d[1] - d[6];
  const int tmp2 = d[2] + d[5], tmp5 = d[2] - d[5], tmp3 = d[3] + d[4], tmp4 = d[3] - d[4];
  const int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
  constexpr int SH = FIRST ? 13 - 2 : 13 + 2;
  if (FIRST) {
    o[0] = (tmp10 + tmp11) << 2;
    o[4] = (tmp10 - tmp11) << 2;
  } else {
    o[0] = descale(tmp10 + tmp11, 2);
    o[4] = descale(tmp10 - tmp11, 2);
  }
  int z1 = (tmp12 + tmp13) * JF_0_541;
  o[2] = descale(z1 + tmp13 * JF_0_765, SH);
  o[6] = descale(z1 + tmp12 * (-JF_1_847), SH);
  z1 = tmp4 + tmp7;
  int z2 = tmp5 + tmp6, z3 = tmp4 + tmp6, z4 = tmp5 + tmp7;
  const int z5 = (z3 + z4) * JF_1_175;
  const int t4 = tmp4 * JF_0_298, t5 = tmp5 * JF_2_053, t6 = tmp6 * JF_3_072, t7 = tmp7 * JF_1_501;
  z1 *= -JF_0_899;
  z2 *= -JF_2_562;
  z3 = z3 * (-JF_1_961) + z5;
  z4 = z4 * (-JF_0_390) + z5;
  o[7] = descale(t4 + z1 + z3, SH);
  o[5] = descale(t5 + z2 + z4, SH);
  o[3] = descale(t6 + z2 + z3, SH);
  o[1] = descale(t7 + z1 + z4, SH);
}

__global__ void __launch_bounds__(kDctThreads)
jpeg_dct_kernel(const vd_frame_t* __restrict__ frames, const __grid_constant__ EncParams P, EncWs ws) {
  __shared__ uint8_t sY[16][kMcusPerCta * 16];
  __shared__ uint8_t sC[2][8][kMcusPerCta * 8];
  __shared__ int sW[kMcusPerCta * 6][64];          // row-pass results
  __shared__ __align__(16) int16_t sOut[kMcusPerCta * 6][64];
  __shared__ uint8_t sZig[64];
  const int tid = threadIdx.x;
  const int f = blockIdx.z, my = blockIdx.y, mx0 = blockIdx.x * kMcusPerCta;
  const int nm = min(kMcusPerCta, P.mcux - mx0);
  const vd_frame_t fr = frames[f];
  const int W = P.W, H = P.H;
  if (tid < 64) sZig[tid] = kZigzagDev[tid];
  // ---- phase 1: one 2x2 pixel quad per task -> 4 Y samples + 1 Cb + 1 Cr ----
  const int real_crows = (H + 1) >> 1;
  for (int q = tid; q < 8 * nm * 8; q += kDctThreads) {
    const int r = q / (nm * 8), c = q - r * (nm * 8);         // chroma row / column inside the tile
    const int cg = mx0 * 8 + c, rg = my * 8 + r;              // ... inside the component
    // luma: pixel (min(y, H-1), min(x, W-1)) — samples of dummy blocks are never used
    int yy[2] = {min(2 * rg, H - 1), min(2 * rg + 1, H - 1)};
    int xx[2] = {min(2 * cg, W - 1), min(2 * cg + 1, W - 1)};
    int cbs = 0, crs = 0;
    // chroma: rows beyond the image are copies of the last DOWN-SAMPLED row (jcprepct.c), whose own second row is the
    // image's last row again when H is odd; columns are extended by the last pixel (expand_right_edge)
    const int rr = min(rg, real_crows - 1);
    const int cy[2] = {2 * rr, min(2 * rr + 1, H - 1)};
    const bool same_rows = (cy[0] == yy[0]) && (cy[1] == yy[1]);
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
      const uint8_t* row = fr.ptr + static_cast<size_t>(yy[dy]) * fr.pitch;
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const uint8_t* p = row + xx[dx] * 3;
        int y, cb, cr;
        rgb_to_ycc(__ldg(p), __ldg(p + 1), __ldg(p + 2), y, cb, cr);
        sY[2 * r + dy][2 * c + dx] = static_cast<uint8_t>(y);
        cbs += cb;
        crs += cr;
      }
    }
    if (!same_rows) {
      cbs = crs = 0;
#pragma unroll
      for (int dy = 0; dy < 2; ++dy) {
        const uint8_t* row = fr.ptr + static_cast<size_t>(cy[dy]) * fr.pitch;
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
          const uint8_t* p = row + xx[dx] * 3;
          int y, cb, cr;
          rgb_to_ycc(__ldg(p), __ldg(p + 1), __ldg(p + 2), y, cb, cr);
          cbs += cb;
          crs += cr;
        }
      }
    }
    const int bias = 1 + (cg & 1);                            // h2v2_downsample: 1, 2, 1, 2, ...
    sC[0][r][c] = static_cast<uint8_t>((cbs + bias) >> 2);
    sC[1][r][c] = static_cast<uint8_t>((crs + bias) >> 2);
  }
  __syncthreads();
  // ---- phase 2: row pass, one task per (block, row) ----
  // block b of MCU m: 0..3 = Y (row b >> 1, column b & 1), 4 = Cb, 5 = Cr
  for (int t = tid; t < nm * 6 * 8; t += kDctThreads) {
    const int blk = t >> 3, row = t & 7;
    const int m = blk / 6, b = blk - m * 6;
    int d[8], o[8];
    if (b < 4) {
      const uint8_t* s = &sY[(b >> 1) * 8 + row][m * 16 + (b & 1) * 8];
#pragma unroll
      for (int i = 0; i < 8; ++i) d[i] = static_cast<int>(s[i]) - 128;
    } else {
      const uint8_t* s = &sC[b - 4][row][m * 8];
#pragma unroll
      for (int i = 0; i < 8; ++i) d[i] = static_cast<int>(s[i]) - 128;
    }
    fdct8<true>(d, o);
#pragma unroll
    for (int i = 0; i < 8; ++i) sW[blk][row * 8 + i] = o[i];
  }
  __syncthreads();
  // ---- phase 3: column pass + quantisation, one task per (block, column) ----
  for (int t = tid; t < nm * 6 * 8; t += kDctThreads) {
    const int blk = t >> 3, col = t & 7;
    const int b = blk % 6;
    int d[8], o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) d[i] = sW[blk][i * 8 + col];
    fdct8<false>(d, o);
    const int tb = b < 4 ? 0 : 1;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int k = i * 8 + col;                              // natural index
      const int v = o[i];
      const unsigned int a = static_cast<unsigned int>(v < 0 ? -v : v) + (P.div[tb][k] >> 1);
      const int qv = static_cast<int>((static_cast<unsigned long long>(a) * P.magic[tb][k]) >> 40);
      sW[blk][k] = v < 0 ? -qv : qv;                          // (each task rewrites only its own column)
    }
  }
  __syncthreads();
  // ---- phase 4: zig-zag, dummy blocks, coalesced store ----
  for (int t = tid; t < nm * 6 * 64; t += kDctThreads) {
    const int blk = t >> 6, k = t & 63;
    sOut[blk][k] = static_cast<int16_t>(sW[blk][sZig[k]]);
  }
  __syncthreads();
  if (tid < nm) {
    // jccoefct.c: luma blocks outside the image inside the last MCU row / column: zero AC, DC of the previous block in
    // MCU order (right edge: the block to the left; a whole dummy block row: the block before the row's first one)
    const int m = tid;
    const int gx = (mx0 + m) * 2, gy = my * 2;
    bool dummy[4];
#pragma unroll
    for (int b = 0; b < 4; ++b) dummy[b] = (gy + (b >> 1) >= P.byY) || (gx + (b & 1) >= P.bxY);
#pragma unroll
    for (int b = 1; b < 4; ++b) {
      if (!dummy[b]) continue;
      const bool row_dummy = gy + (b >> 1) >= P.byY;
      const int ref = row_dummy ? (b & ~1) - 1 : b - 1;       // row 1 dummy: block 1; right edge: left neighbour
      const int16_t dc = sOut[m * 6 + ref][0];
      for (int k = 1; k < 64; ++k) sOut[m * 6 + b][k] = 0;
      sOut[m * 6 + b][0] = dc;
    }
  }
  __syncthreads();
  {
    const size_t base = (static_cast<size_t>(f) * (P.mcux * P.mcuy) + static_cast<size_t>(my) * P.mcux + mx0) * 6 * 64;
    uint4* dst = reinterpret_cast<uint4*>(ws.coef + base);
    const uint4* src = reinterpret_cast<const uint4*>(&sOut[0][0]);
    for (int i = tid; i < nm * 6 * 8; i += kDctThreads) dst[i] = src[i];
  }
}

// ---------------------------------------------------------------------------------------------
// 2./3. Huffman coding: count, then emit
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int nbits_of(int v) { return 32 - __clz(v); }      // v >= 0

// DC predictor of block `blk` (frame-local index = mcu * 6 + b): previous block of the same component in scan order
__device__ __forceinline__ int prev_dc(const int16_t* __restrict__ coef, int blk) {
  const int m = blk / 6, b = blk - m * 6;
  if (b >= 1 && b <= 3) return coef[static_cast<size_t>(blk - 1) * 64];
  if (m == 0) return 0;
  return coef[(static_cast<size_t>(m - 1) * 6 + (b == 0 ? 3 : b)) * 64];
}

// walks one block; EMIT = false: returns its bit count; EMIT = true: writes the code words at bit position `pos`
template <bool EMIT>
__device__ __forceinline__ uint32_t code_block(const int16_t* __restrict__ blk, int pred, const uint32_t* __restrict__ dc_code,
                                               const uint32_t* __restrict__ ac_code, uint32_t* __restrict__ words,
                                               unsigned long long pos) {
  uint32_t bits = 0;
  unsigned long long acc = 0;     // EMIT: pending bits, left-aligned below `fill`
  int fill = 0;
  auto put = [&](uint32_t code, int len) {
    bits += len;
    if (EMIT && len) {
      acc = (acc << len) | code;
      fill += len;
      if (fill >= 32) {           // flush one 32-bit word's worth at the current bit position
        const int rest = fill - 32;
        const uint32_t w = static_cast<uint32_t>(acc >> rest);
        const uint32_t sh = static_cast<uint32_t>(pos & 31);
        uint32_t* wp = words + (pos >> 5);
        atomicOr(wp, w >> sh);
        if (sh) atomicOr(wp + 1, w << (32 - sh));
        pos += 32;
        fill = rest;
        acc &= (1ull << rest) - 1ull;
      }
    }
  };
  // 16-byte loads of the 64 coefficients
  const uint4* b4 = reinterpret_cast<const uint4*>(blk);
  int diff = static_cast<int>(blk[0]) - pred;
  {
    int t = diff, t2 = diff;
    if (t < 0) { t = -t; t2 = diff - 1; }
    const int n = nbits_of(t);
    const uint32_t e = dc_code[n];
    put(e >> 8, e & 0xFF);
    put(static_cast<uint32_t>(t2) & ((1u << n) - 1u), n);
  }
  int run = 0;
  for (int g = 0; g < 8; ++g) {
    const uint4 v4 = b4[g];
    const uint32_t wv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (g == 0 && i == 0) continue;
      const int v = static_cast<int16_t>((wv[i >> 1] >> ((i & 1) * 16)) & 0xFFFF);
      if (v == 0) {
        ++run;
        continue;
      }
      while (run > 15) {
        const uint32_t e = ac_code[0xF0];
        put(e >> 8, e & 0xFF);
        run -= 16;
      }
      int t = v, t2 = v;
      if (t < 0) { t = -t; t2 = v - 1; }
      const int n = nbits_of(t);
      const uint32_t e = ac_code[(run << 4) + n];
      put(e >> 8, e & 0xFF);
      put(static_cast<uint32_t>(t2) & ((1u << n) - 1u), n);
      run = 0;
    }
  }
  if (run > 0) {
    const uint32_t e = ac_code[0];
    put(e >> 8, e & 0xFF);
  }
  if (EMIT && fill > 0) {         // the tail (< 32 bits), left-aligned
    const uint32_t w = static_cast<uint32_t>(acc << (32 - fill));
    const uint32_t sh = static_cast<uint32_t>(pos & 31);
    uint32_t* wp = words + (pos >> 5);
    atomicOr(wp, w >> sh);
    if (sh && fill > static_cast<int>(32 - sh)) atomicOr(wp + 1, w << (32 - sh));
  }
  return bits;
}

__global__ void __launch_bounds__(kBlocksPerCta)
jpeg_count_kernel(const __grid_constant__ EncParams P, EncWs ws) {
  __shared__ uint32_t s_warp[kBlocksPerCta / 32];
  __shared__ uint32_t s_dc[2][12], s_ac[2][256];
  const int tid = threadIdx.x, f = blockIdx.y;
  for (int i = tid; i < 24; i += kBlocksPerCta) s_dc[i / 12][i % 12] = P.dc_code[i / 12][i % 12];
  for (int i = tid; i < 512; i += kBlocksPerCta) s_ac[i >> 8][i & 255] = P.ac_code[i >> 8][i & 255];
  __syncthreads();
  const int blk = blockIdx.x * kBlocksPerCta + tid;
  const int16_t* coef = ws.coef + static_cast<size_t>(f) * ws.nblk * 64;
  uint32_t bits = 0;
  if (blk < ws.nblk) {
    const int tb = (blk % 6) < 4 ? 0 : 1;
    bits = code_block<false>(coef + static_cast<size_t>(blk) * 64, prev_dc(coef, blk), s_dc[tb], s_ac[tb], nullptr, 0);
  }
  // exclusive scan inside the CTA
  uint32_t incl = bits;
  const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t n = __shfl_up_sync(0xFFFFFFFFu, incl, o);
    if (lane >= o) incl += n;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  uint32_t wbase = 0;
  for (int w = 0; w < warp; ++w) wbase += s_warp[w];
  ws.blk_off[(static_cast<size_t>(f) * ws.ctas_per_frame + blockIdx.x) * kBlocksPerCta + tid] = wbase + incl - bits;
  if (tid == kBlocksPerCta - 1) ws.cta_bits[static_cast<size_t>(f) * ws.ctas_per_frame + blockIdx.x] = wbase + incl;
}

__global__ void __launch_bounds__(kBlocksPerCta)
jpeg_emit_kernel(const __grid_constant__ EncParams P, EncWs ws, uint32_t* __restrict__ status) {
  __shared__ uint32_t s_dc[2][12], s_ac[2][256];
  __shared__ unsigned long long s_base;
  __shared__ unsigned long long s_part[kBlocksPerCta / 32];
  const int tid = threadIdx.x, f = blockIdx.y;
  for (int i = tid; i < 24; i += kBlocksPerCta) s_dc[i / 12][i % 12] = P.dc_code[i / 12][i % 12];
  for (int i = tid; i < 512; i += kBlocksPerCta) s_ac[i >> 8][i & 255] = P.ac_code[i >> 8][i & 255];
  // bit offset of this CTA = sum of the totals of the CTAs before it (a few hundred values)
  const uint32_t* cb = ws.cta_bits + static_cast<size_t>(f) * ws.ctas_per_frame;
  unsigned long long part = 0, all = 0;
  for (int i = tid; i < ws.ctas_per_frame; i += kBlocksPerCta) {
    const unsigned long long v = cb[i];
    all += v;
    if (i < static_cast<int>(blockIdx.x)) part += v;
  }
  // reduce `part` (and `all`, needed by CTA 0 only) over the CTA
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    part += __shfl_down_sync(0xFFFFFFFFu, part, o);
    all += __shfl_down_sync(0xFFFFFFFFu, all, o);
  }
  __shared__ unsigned long long s_all[kBlocksPerCta / 32];
  if ((tid & 31) == 0) { s_part[tid >> 5] = part; s_all[tid >> 5] = all; }
  __syncthreads();
  if (tid == 0) {
    unsigned long long p = 0, a = 0;
    for (int w = 0; w < kBlocksPerCta / 32; ++w) { p += s_part[w]; a += s_all[w]; }
    s_base = p;
    if (blockIdx.x == 0) {
      const bool fits = a <= static_cast<unsigned long long>(ws.cap_words) * 32ull - 64ull;
      ws.total_bits[f] = fits ? static_cast<uint32_t>(a) : 0xFFFFFFFFu;
      if (!fits) atomicOr(status, VD_STAT_JPEG_OVERFLOW);
    }
    s_all[0] = a;
  }
  __syncthreads();
  if (s_all[0] > static_cast<unsigned long long>(ws.cap_words) * 32ull - 64ull) return;     // frame does not fit: nothing is written
  const int blk = blockIdx.x * kBlocksPerCta + tid;
  if (blk >= ws.nblk) return;
  const int16_t* coef = ws.coef + static_cast<size_t>(f) * ws.nblk * 64;
  const unsigned long long pos = s_base + ws.blk_off[(static_cast<size_t>(f) * ws.ctas_per_frame + blockIdx.x) * kBlocksPerCta + tid];
  const int tb = (blk % 6) < 4 ? 0 : 1;
  code_block<true>(coef + static_cast<size_t>(blk) * 64, prev_dc(coef, blk), s_dc[tb], s_ac[tb],
                   ws.words + static_cast<size_t>(f) * ws.cap_words, pos);
}

// ---------------------------------------------------------------------------------------------
// 4./5. byte stuffing + header + EOI
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t stream_byte(const uint32_t* __restrict__ words, uint32_t i, uint32_t nbytes, uint32_t total_bits) {
  uint32_t b = (words[i >> 2] >> (24 - 8 * (i & 3))) & 0xFFu;
  if (i == nbytes - 1 && (total_bits & 7u)) b |= (1u << (8 - (total_bits & 7u))) - 1u;      // pad the last byte with one-bits
  return b;
}

__global__ void __launch_bounds__(256)
jpeg_ffcount_kernel(EncWs ws) {
  const int f = blockIdx.y;
  const uint32_t total_bits = ws.total_bits[f];
  if (total_bits == 0xFFFFFFFFu) return;
  const uint32_t nbytes = (total_bits + 7) >> 3;
  const uint32_t c0 = blockIdx.x * kChunk;
  if (c0 >= nbytes) return;
  const uint32_t* words = ws.words + static_cast<size_t>(f) * ws.cap_words;
  uint32_t n = 0;
  for (uint32_t i = c0 + threadIdx.x; i < min(c0 + kChunk, nbytes); i += 256) n += stream_byte(words, i, nbytes, total_bits) == 0xFFu;
  n = __reduce_add_sync(0xFFFFFFFFu, n);
  __shared__ uint32_t s[8];
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = n;
  __syncthreads();
  if (threadIdx.x == 0) ws.ff_count[static_cast<size_t>(f) * ws.chunks_cap + blockIdx.x] = s[0] + s[1] + s[2] + s[3] + s[4] + s[5] + s[6] + s[7];
}

__global__ void __launch_bounds__(256)
jpeg_finish_kernel(const __grid_constant__ EncParams P, EncWs ws, uint8_t* __restrict__ out, size_t out_pitch,
                   int32_t* __restrict__ out_sizes, uint32_t* __restrict__ status) {
  const int f = blockIdx.y, tid = threadIdx.x;
  const uint32_t total_bits = ws.total_bits[f];
  uint8_t* dst = out + static_cast<size_t>(f) * out_pitch;
  if (total_bits == 0xFFFFFFFFu) {
    if (blockIdx.x == 0 && tid == 0) out_sizes[f] = -1;
    return;
  }
  const uint32_t nbytes = (total_bits + 7) >> 3;
  const uint32_t nchunks = (nbytes + kChunk - 1) / kChunk;
  if (blockIdx.x > nchunks) return;
  // stuffed bytes before this chunk (and in total)
  const uint32_t* fc = ws.ff_count + static_cast<size_t>(f) * ws.chunks_cap;
  uint32_t before = 0, all = 0;
  for (uint32_t i = tid; i < nchunks; i += 256) {
    const uint32_t v = fc[i];
    all += v;
    if (i < blockIdx.x) before += v;
  }
  before = __reduce_add_sync(0xFFFFFFFFu, before);
  all = __reduce_add_sync(0xFFFFFFFFu, all);
  __shared__ uint32_t sb[8], sa[8], s_scan[8];
  if ((tid & 31) == 0) { sb[tid >> 5] = before; sa[tid >> 5] = all; }
  __syncthreads();
  before = sb[0] + sb[1] + sb[2] + sb[3] + sb[4] + sb[5] + sb[6] + sb[7];
  all = sa[0] + sa[1] + sa[2] + sa[3] + sa[4] + sa[5] + sa[6] + sa[7];
  const size_t file_size = static_cast<size_t>(P.header_len) + nbytes + all + 2;
  const bool fits = file_size <= out_pitch;
  if (blockIdx.x == nchunks) {                        // one extra CTA: header, EOI, size
    if (fits) {
      for (int i = tid; i < P.header_len; i += 256) dst[i] = P.header[i];
      if (tid == 0) {
        dst[file_size - 2] = 0xFF;
        dst[file_size - 1] = 0xD9;
      }
    }
    if (tid == 0) {
      out_sizes[f] = fits ? static_cast<int32_t>(file_size) : -1;
      if (!fits) atomicOr(status, VD_STAT_JPEG_OVERFLOW);
    }
    return;
  }
  if (!fits) return;
  const uint32_t* words = ws.words + static_cast<size_t>(f) * ws.cap_words;
  const uint32_t c0 = blockIdx.x * kChunk;
  // each thread owns 4 consecutive bytes; exclusive scan of its 0xFF count over the CTA
  uint32_t b[4], nff = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint32_t i = c0 + 4 * tid + k;
    b[k] = i < nbytes ? stream_byte(words, i, nbytes, total_bits) : 0x100u;
    nff += b[k] == 0xFFu;
  }
  uint32_t incl = nff;
  const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t n = __shfl_up_sync(0xFFFFFFFFu, incl, o);
    if (lane >= o) incl += n;
  }
  if (lane == 31) s_scan[warp] = incl;
  __syncthreads();
  uint32_t wbase = 0;
  for (int w = 0; w < warp; ++w) wbase += s_scan[w];
  size_t o = static_cast<size_t>(P.header_len) + c0 + before + 4 * tid + (wbase + incl - nff);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (b[k] == 0x100u) break;
    dst[o++] = static_cast<uint8_t>(b[k]);
    if (b[k] == 0xFFu) dst[o++] = 0;
  }
}

// ---------------------------------------------------------------------------------------------
// host: tables and header for (W, H, quality)
// ---------------------------------------------------------------------------------------------
void make_codes(const uint8_t* bits, const uint8_t* vals, uint32_t* table, bool ac) {
  uint32_t code = 0;
  int k = 0;
  for (int len = 1; len <= 16; ++len) {
    for (int i = 0; i < bits[len - 1]; ++i, ++k, ++code) table[vals[k]] = (code << 8) | static_cast<uint32_t>(len);
    code <<= 1;
  }
  (void)ac;
}

int put_seg(uint8_t* o, int pos, int marker, const uint8_t* payload, int n) {
  o[pos++] = 0xFF;
  o[pos++] = static_cast<uint8_t>(marker);
  o[pos++] = static_cast<uint8_t>((n + 2) >> 8);
  o[pos++] = static_cast<uint8_t>((n + 2) & 0xFF);
  memcpy(o + pos, payload, static_cast<size_t>(n));
  return pos + n;
}

void make_params(int W, int H, int quality, EncParams* P) {
  memset(P, 0, sizeof(*P));
  P->W = W; P->H = H;
  P->mcux = vd_cdiv(W, 16); P->mcuy = vd_cdiv(H, 16);
  P->bxY = vd_cdiv(W, 8); P->byY = vd_cdiv(H, 8);
  if (quality < 1) quality = 1;
  if (quality > 100) quality = 100;
  const int scale = quality < 50 ? 5000 / quality : 200 - 2 * quality;      // jcparam.c jpeg_quality_scaling
  uint8_t q[2][64];
  for (int t = 0; t < 2; ++t)
    for (int k = 0; k < 64; ++k) {
      long v = (static_cast<long>(t ? kStdChromaQ[k] : kStdLumaQ[k]) * scale + 50) / 100;
      v = v < 1 ? 1 : (v > 255 ? 255 : v);                                    // force_baseline
      q[t][k] = static_cast<uint8_t>(v);
      P->div[t][k] = static_cast<uint16_t>(v << 3);
      P->magic[t][k] = ((1ull << 40) + static_cast<unsigned long long>(v << 3) - 1ull) / static_cast<unsigned long long>(v << 3);
    }
  make_codes(kDcLumaBits, kDcVals, P->dc_code[0], false);
  make_codes(kDcChromaBits, kDcVals, P->dc_code[1], false);
  make_codes(kAcLumaBits, kAcLumaVals, P->ac_code[0], true);
  make_codes(kAcChromaBits, kAcChromaVals, P->ac_code[1], true);
  // jcmarker.c: SOI, APP0 (JFIF 1.01, no units, 1:1), DQT x2, SOF0, DHT x4, SOS
  uint8_t* o = P->header;
  int pos = 0;
  o[pos++] = 0xFF; o[pos++] = 0xD8;
  const uint8_t app0[14] = {'J', 'F', 'I', 'F', 0, 1, 1, 0, 0, 1, 0, 1, 0, 0};
  pos = put_seg(o, pos, 0xE0, app0, 14);
  for (int t = 0; t < 2; ++t) {
    uint8_t dqt[65];
    dqt[0] = static_cast<uint8_t>(t);
    for (int k = 0; k < 64; ++k) dqt[1 + k] = q[t][kZigzag[k]];
    pos = put_seg(o, pos, 0xDB, dqt, 65);
  }
  const uint8_t sof[15] = {8, static_cast<uint8_t>(H >> 8), static_cast<uint8_t>(H & 0xFF), static_cast<uint8_t>(W >> 8),
                           static_cast<uint8_t>(W & 0xFF), 3, 1, 0x22, 0, 2, 0x11, 1, 3, 0x11, 1};
  pos = put_seg(o, pos, 0xC0, sof, 15);
  struct { int id; const uint8_t* bits; const uint8_t* vals; int n; } dht[4] = {
      {0x00, kDcLumaBits, kDcVals, 12}, {0x10, kAcLumaBits, kAcLumaVals, 162},
      {0x01, kDcChromaBits, kDcVals, 12}, {0x11, kAcChromaBits, kAcChromaVals, 162}};
  for (int t = 0; t < 4; ++t) {
    uint8_t seg[1 + 16 + 162];
    seg[0] = static_cast<uint8_t>(dht[t].id);
    memcpy(seg + 1, dht[t].bits, 16);
    memcpy(seg + 17, dht[t].vals, static_cast<size_t>(dht[t].n));
    pos = put_seg(o, pos, 0xC4, seg, 17 + dht[t].n);
  }
  const uint8_t sos[10] = {3, 1, 0x00, 2, 0x11, 3, 0x11, 0, 63, 0};
  pos = put_seg(o, pos, 0xDA, sos, 10);
  P->header_len = pos;
}

}  // namespace

extern "C" size_t vd_jpeg_encode_workspace_bytes(int B, int H, int W) {
  if (B <= 0 || H <= 0 || W <= 0) return 0;
  const size_t nblk = static_cast<size_t>(vd_cdiv(W, 16)) * vd_cdiv(H, 16) * 6;
  const size_t ctas = (nblk + kBlocksPerCta - 1) / kBlocksPerCta;
  // word stream: 2 bytes per pixel per frame is more than any photographic frame needs at quality <= 95 (a 1080p frame
  // at 95 is ~0.9 MB = 0.45 B/px); a frame that does not fit is reported (VD_STAT_JPEG_OVERFLOW, size -1), not truncated
  const size_t stream = (static_cast<size_t>(H) * W * 2 + 65536) & ~static_cast<size_t>(kChunk - 1);
  return al256(static_cast<size_t>(B) * nblk * 128) + al256(static_cast<size_t>(B) * ctas * kBlocksPerCta * 4) +
         al256(static_cast<size_t>(B) * ctas * 4) + al256(static_cast<size_t>(B) * 4) +
         static_cast<size_t>(B) * (stream + stream / kChunk * 4 + 64) + 8192;
}

extern "C" int vd_jpeg_encode_batch(const vd_frame_t* frames, int B, int H, int W, int quality, uint8_t* out,
                                    size_t out_pitch, int32_t* out_sizes, uint32_t* status, void* workspace,
                                    size_t workspace_bytes, vd_stream_t stream) {
  VD_REQUIRE(frames && out && out_sizes && status && workspace, "vd_jpeg_encode_batch: null argument");
  VD_REQUIRE(B > 0 && B <= 65535 && H > 0 && W > 0 && H <= 65535 && W <= 65535, "vd_jpeg_encode_batch: bad extent");
  VD_REQUIRE(quality >= 1 && quality <= 100, "vd_jpeg_encode_batch: quality must be in [1,100]");
  VD_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "vd_jpeg_encode_batch: workspace must be 256-B aligned");
  VD_REQUIRE(workspace_bytes >= vd_jpeg_encode_workspace_bytes(B, H, W), "vd_jpeg_encode_batch: workspace too small");
  VD_REQUIRE(vd_cdiv(H, 16) <= 65535, "vd_jpeg_encode_batch: frame too tall");
  static_assert(sizeof(EncParams) <= 8192, "EncParams is passed by value (CUDA >= 12.1: up to 32 KB of kernel parameters)");
  EncParams P;
  make_params(W, H, quality, &P);
  size_t zero_from = 0, zero_bytes = 0;
  const EncWs ws = carve(workspace, B, H, W, workspace_bytes, &zero_from, &zero_bytes);
  VD_REQUIRE(ws.cap_words >= 1024, "vd_jpeg_encode_batch: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  VD_CHECK_CUDA(cudaMemsetAsync(static_cast<uint8_t*>(workspace) + zero_from, 0, zero_bytes, st));
  jpeg_dct_kernel<<<dim3(vd_cdiv(P.mcux, kMcusPerCta), P.mcuy, B), kDctThreads, 0, st>>>(frames, P, ws);
  VD_LAUNCH_CHECK();
  jpeg_count_kernel<<<dim3(ws.ctas_per_frame, B), kBlocksPerCta, 0, st>>>(P, ws);
  VD_LAUNCH_CHECK();
  jpeg_emit_kernel<<<dim3(ws.ctas_per_frame, B), kBlocksPerCta, 0, st>>>(P, ws, status);
  VD_LAUNCH_CHECK();
  jpeg_ffcount_kernel<<<dim3(ws.chunks_cap, B), 256, 0, st>>>(ws);
  VD_LAUNCH_CHECK();
  jpeg_finish_kernel<<<dim3(ws.chunks_cap + 1, B), 256, 0, st>>>(P, ws, out, out_pitch, out_sizes, status);
  VD_LAUNCH_CHECK();
  return VD_OK;
}
