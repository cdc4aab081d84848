// tiff_lzw.h — baseline TIFF 6.0 writer with LZW compression (host only, multi-threaded over strips).
// Replaces tifffile.imwrite(path, array, compression='lzw') of the reference's outputs
// (reconstruct_full_images.py:724-734: original_image.tif, prediction_mask.tif, ground_truth_mask.tif;
// segmentation_inference.py:455-464: masks / probability maps).  A 32768^2 mosaic is 1-3 GB of pixels: the
// strips are compressed independently (TIFF LZW restarts its string table per strip), one std::thread per
// range of strips, then written in order.  Readable by libtiff (cv2.imread, PIL) and tifffile.
//
// Encoder = the algorithm of TIFF 6.0 section 13 with libtiff's code-width schedule ("early change"): codes
// are packed MSB first, start 9 bits wide, ClearCode 256, EndOfInformation 257, first table entry 258; the
// width grows when the next free entry exceeds 2^width - 1 and the table is cleared when entry 4093 has been
// assigned.  Uncompressed sample order is the caller's (the reference hands tifffile a BGR array for
// original_image.tif, so that is what lands in the file).
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <thread>
#include <vector>

namespace adp_tiff {

// LZW of one strip.  Hash table: key = (prefix code << 8) | byte, open addressing, cleared by a generation stamp.
struct LzwScratch {
  std::vector<uint32_t> key, gen;
  std::vector<uint16_t> val;
  uint32_t cur = 0;
  LzwScratch() : key(1 << 14), gen(1 << 14, 0u), val(1 << 14) {}
};

inline void lzw_encode(const uint8_t *src, size_t n, std::vector<uint8_t> &out, LzwScratch &sc) {
  constexpr int kClear = 256, kEoi = 257, kFirst = 258, kCodeMax = 4095, kHashMask = (1 << 14) - 1;
  uint32_t *hkey = sc.key.data(), *hgen = sc.gen.data();
  uint16_t *hval = sc.val.data();
  uint32_t gen = ++sc.cur;
  // worst case: every input byte emits one 12-bit code, plus a clear code every 3837 entries
  out.resize(n + n / 2 + n / 1024 + 64);
  uint8_t *o = out.data();
  uint64_t acc = 0;
  int nacc = 0;
  auto put = [&](uint32_t code, int nbits) {
    acc = (acc << nbits) | code;
    nacc += nbits;
    while (nacc >= 8) { *o++ = (uint8_t)(acc >> (nacc - 8)); nacc -= 8; }
  };
  int nbits = 9, maxcode = 511, free_ent = kFirst;
  put(kClear, nbits);
  if (n != 0) {
    uint32_t ent = src[0];
    for (size_t i = 1; i < n; ++i) {
      const uint32_t c = src[i];
      const uint32_t key = (ent << 8) | c;
      uint32_t h = (key * 2654435761u) >> 18;          // 14 bits
      bool found = false;
      while (hgen[h] == gen) {
        if (hkey[h] == key) { ent = hval[h]; found = true; break; }
        h = (h + 1) & kHashMask;
      }
      if (found) continue;
      put(ent, nbits);
      hgen[h] = gen; hkey[h] = key; hval[h] = (uint16_t)free_ent;
      ++free_ent;
      ent = c;
      if (free_ent == kCodeMax - 1) {                  // table full: clear and restart at 9 bits
        put(kClear, nbits);
        gen = ++sc.cur;
        free_ent = kFirst; nbits = 9; maxcode = 511;
      } else if (free_ent > maxcode) {
        ++nbits; maxcode = (1 << nbits) - 1;
      }
    }
    put(ent, nbits);
    ++free_ent;                                        // the last code counts towards the width schedule (the decoder adds an entry)
    if (free_ent == kCodeMax - 1) { put(kClear, nbits); nbits = 9; }
    else if (free_ent > maxcode) ++nbits;
  }
  put(kEoi, nbits);
  if (nacc > 0) *o++ = (uint8_t)(acc << (8 - nacc));
  out.resize((size_t)(o - out.data()));
}

inline void put16(std::vector<uint8_t> &b, uint16_t v) { b.push_back((uint8_t)(v & 255)); b.push_back((uint8_t)(v >> 8)); }
inline void put32(std::vector<uint8_t> &b, uint32_t v) { put16(b, (uint16_t)(v & 0xFFFF)); put16(b, (uint16_t)(v >> 16)); }

// data: H x W x channels uint8, row-major, samples interleaved.  Returns 0, or a negative code with msg filled.
inline int write_lzw(const char *path, const uint8_t *data, int64_t H, int64_t W, int channels, int rows_per_strip, int threads,
                     std::string &msg) {
  if (!path || !data || H <= 0 || W <= 0 || (channels != 1 && channels != 3)) { msg = "tiff: bad arguments"; return -1; }
  const size_t row_bytes = (size_t)W * channels;
  if (rows_per_strip <= 0) {
    rows_per_strip = (int)((1u << 20) / row_bytes);  // about 1 MB of pixels per strip
    if (rows_per_strip < 1) rows_per_strip = 1;
  }
  if (rows_per_strip > H) rows_per_strip = (int)H;
  const int64_t nstrips = (H + rows_per_strip - 1) / rows_per_strip;
  if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
  if (threads < 1) threads = 1;
  if (threads > nstrips) threads = (int)nstrips;
  std::vector<std::vector<uint8_t>> comp((size_t)nstrips);
  auto work = [&](int t) {
    LzwScratch sc;
    for (int64_t s = t; s < nstrips; s += threads) {
      const int64_t r0 = s * rows_per_strip;
      const int64_t nr = (r0 + rows_per_strip <= H) ? rows_per_strip : (H - r0);
      lzw_encode(data + (size_t)r0 * row_bytes, (size_t)nr * row_bytes, comp[(size_t)s], sc);
      comp[(size_t)s].shrink_to_fit();
    }
  };
  if (threads == 1) work(0);
  else {
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) pool.emplace_back(work, t);
    for (auto &th : pool) th.join();
  }
  uint64_t total = 8;
  for (auto &c : comp) total += c.size() + (c.size() & 1);
  const uint64_t ifd_bytes = 2 + 10 * 12 + 4 + (uint64_t)nstrips * 8 + 6;
  if (total + ifd_bytes >= 0xFFFFFFF0ull) { msg = "tiff: compressed image exceeds the 4 GB limit of classic TIFF"; return -2; }
  FILE *f = fopen(path, "wb");
  if (!f) { msg = std::string("tiff: cannot open ") + path; return -3; }
  std::vector<uint8_t> hdr;
  hdr.push_back('I'); hdr.push_back('I'); put16(hdr, 42);
  put32(hdr, (uint32_t)total);                       // offset of the IFD (after the strips)
  bool ok = fwrite(hdr.data(), 1, hdr.size(), f) == hdr.size();
  std::vector<uint32_t> offs((size_t)nstrips), cnts((size_t)nstrips);
  uint32_t pos = 8;
  const uint8_t zero = 0;
  for (int64_t s = 0; s < nstrips && ok; ++s) {
    const auto &c = comp[(size_t)s];
    offs[(size_t)s] = pos; cnts[(size_t)s] = (uint32_t)c.size();
    ok = fwrite(c.data(), 1, c.size(), f) == c.size();
    pos += (uint32_t)c.size();
    if (c.size() & 1) { ok = ok && fwrite(&zero, 1, 1, f) == 1; ++pos; }   // word alignment
  }
  // IFD: 10 entries, then the out-of-line arrays
  std::vector<uint8_t> ifd;
  const uint32_t ifd_off = pos;
  const uint32_t arrays_off = ifd_off + 2 + 10 * 12 + 4;
  uint32_t bps_off = 0, so_off = 0, sc_off = 0, cur = arrays_off;
  if (channels == 3) { bps_off = cur; cur += 6; }
  if (nstrips > 1) { so_off = cur; cur += (uint32_t)nstrips * 4; sc_off = cur; cur += (uint32_t)nstrips * 4; }
  auto entry = [&](uint16_t tag, uint16_t type, uint32_t count, uint32_t value) { put16(ifd, tag); put16(ifd, type); put32(ifd, count); put32(ifd, value); };
  put16(ifd, 10);
  entry(256, 4, 1, (uint32_t)W);                                              // ImageWidth
  entry(257, 4, 1, (uint32_t)H);                                              // ImageLength
  if (channels == 3) entry(258, 3, 3, bps_off); else entry(258, 3, 1, 8);     // BitsPerSample
  entry(259, 3, 1, 5);                                                        // Compression = LZW
  entry(262, 3, 1, channels == 3 ? 2 : 1);                                    // Photometric: RGB / BlackIsZero
  if (nstrips > 1) entry(273, 4, (uint32_t)nstrips, so_off); else entry(273, 4, 1, offs[0]);   // StripOffsets
  entry(277, 3, 1, (uint32_t)channels);                                       // SamplesPerPixel
  entry(278, 4, 1, (uint32_t)rows_per_strip);                                 // RowsPerStrip
  if (nstrips > 1) entry(279, 4, (uint32_t)nstrips, sc_off); else entry(279, 4, 1, cnts[0]);   // StripByteCounts
  entry(284, 3, 1, 1);                                                        // PlanarConfiguration = chunky
  put32(ifd, 0);                                                              // no next IFD
  if (channels == 3) { put16(ifd, 8); put16(ifd, 8); put16(ifd, 8); }
  if (nstrips > 1) {
    for (auto v : offs) put32(ifd, v);
    for (auto v : cnts) put32(ifd, v);
  }
  ok = ok && fwrite(ifd.data(), 1, ifd.size(), f) == ifd.size();
  ok = (fclose(f) == 0) && ok;
  if (!ok) { msg = std::string("tiff: write error on ") + path; return -4; }
  return 0;
}

}  // namespace adp_tiff
