// Self-describing stream framing (host side, C++): preamble + protobuf header.
//
//   reference writer  src/mgard-x/Metadata/Metadata.cpp:249-462 (MetadataBase::Serialize)
//   reference reader  src/mgard-x/Metadata/Metadata.cpp:475-739, CPU reader src/format.cpp:202-525
//   schema            src/mgard.proto:1-185
//
// The header is the proto3 canonical serialisation of mgard.pb.Header (fields
// in number order, default values omitted, repeated scalars packed) — written
// by hand here (varint / fixed64 / length-delimited only) so that no protobuf
// runtime is needed.  The preamble is "MGARD" | u64 header size | u32
// CRC32(header): little-endian in MGARD-X streams (Metadata.cpp:441-459, native
// byte order), BIG-endian in MGARD-CPU streams (serialize / deserialize of
// include/format.tpp:11-41, used by write_metadata / read_metadata,
// src/format.cpp:202-233).
#include <cmath>
#include <cstring>
#include <string>

#include "format.h"

namespace {

// CRC-32 (zlib polynomial 0xEDB88320), as crc32_z in Metadata.cpp:34-36
struct Crc32Table {
  uint32_t t[256];
  Crc32Table() {
    for (uint32_t i = 0; i < 256; i++) {
      uint32_t c = i;
      for (int k = 0; k < 8; k++)
        c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
      t[i] = c;
    }
  }
};
uint32_t crc32_bytes(const uint8_t *p, size_t n) {
  static const Crc32Table table; // thread-safe initialisation (C++11 magic static)
  uint32_t c = 0xFFFFFFFFu;
  for (size_t i = 0; i < n; i++)
    c = table.t[(c ^ p[i]) & 0xFF] ^ (c >> 8);
  return c ^ 0xFFFFFFFFu;
}

void put_varint(std::string &o, uint64_t x) {
  while (x >= 0x80) {
    o.push_back((char)((x & 0x7F) | 0x80));
    x >>= 7;
  }
  o.push_back((char)x);
}
void f_varint(std::string &o, int field, uint64_t x) {
  if (x == 0)
    return; // proto3: default values are not serialised
  put_varint(o, (uint64_t)field << 3);
  put_varint(o, x);
}
void f_double(std::string &o, int field, double x) {
  uint64_t bits;
  memcpy(&bits, &x, 8);
  if (bits == 0)
    return;
  put_varint(o, ((uint64_t)field << 3) | 1);
  o.append((const char *)&bits, 8);
}
void f_msg(std::string &o, int field, const std::string &body) {
  put_varint(o, ((uint64_t)field << 3) | 2);
  put_varint(o, body.size());
  o += body;
}

struct Reader {
  const uint8_t *p, *end;
  bool ok = true;
  uint64_t varint() {
    uint64_t x = 0;
    int shift = 0;
    while (p < end && shift < 64) {
      uint8_t b = *p++;
      x |= (uint64_t)(b & 0x7F) << shift;
      if (!(b & 0x80))
        return x;
      shift += 7;
    }
    ok = false;
    return 0;
  }
  // returns false at end; otherwise field number/wire type and a sub-reader or value
  bool next(int &field, int &wt, uint64_t &val, Reader &sub) {
    if (p >= end || !ok)
      return false;
    uint64_t key = varint();
    field = (int)(key >> 3);
    wt = (int)(key & 7);
    switch (wt) {
    case 0: val = varint(); break;
    case 1:
      if (end - p < 8) { ok = false; return false; }
      memcpy(&val, p, 8);
      p += 8;
      break;
    case 2: {
      uint64_t len = varint();
      if ((uint64_t)(end - p) < len) { ok = false; return false; }
      sub.p = p;
      sub.end = p + len;
      sub.ok = true;
      p += len;
      break;
    }
    case 5:
      if (end - p < 4) { ok = false; return false; }
      val = 0;
      memcpy(&val, p, 4);
      p += 4;
      break;
    default: ok = false; return false;
    }
    return ok;
  }
};

double as_double(uint64_t bits) {
  double d;
  memcpy(&d, &bits, 8);
  return d;
}

} // namespace

namespace {
std::vector<uint8_t> with_preamble(const std::string &hdr, bool big_endian) {
  std::vector<uint8_t> out;
  out.insert(out.end(), {'M', 'G', 'A', 'R', 'D'});
  uint64_t hs = hdr.size();
  for (int i = 0; i < 8; i++)
    out.push_back((uint8_t)(hs >> (8 * (big_endian ? 7 - i : i))));
  uint32_t crc = crc32_bytes((const uint8_t *)hdr.data(), hdr.size());
  for (int i = 0; i < 4; i++)
    out.push_back((uint8_t)(crc >> (8 * (big_endian ? 3 - i : i))));
  out.insert(out.end(), hdr.begin(), hdr.end());
  return out;
}

// Header of mgard::compress (reference include/compress.tpp:41-55):
// populate_defaults (src/format.cpp:102-140, build without MGARD_ZSTD),
// TensorMeshHierarchy::populate (include/TensorMeshHierarchy.tpp:293-348) and the
// error-control block.  Sub-messages that the reference touches through
// mutable_*() are present even when empty; domain_decomposition and
// bitplane_encoding are never touched and stay absent.
std::vector<uint8_t> encode_cpu_header(const mgb_header &h) {
  std::string ver, fver;
  f_varint(ver, 1, 1); // MGARD_VERSION 1.6.0 (CMakeLists.txt:13-15)
  f_varint(ver, 2, 6);
  f_varint(fver, 1, 1); // MGARD_FILE_VERSION 1.0.0 (:17-19)
  std::string topo;
  f_varint(topo, 1, (uint64_t)h.ndim);
  {
    std::string packed;
    for (int d = 0; d < h.ndim; d++)
      put_varint(packed, h.shape[d]);
    f_msg(topo, 2, packed);
  }
  std::string dom;
  f_msg(dom, 2, topo);
  if (!h.coords.empty()) {
    f_varint(dom, 3, 1); // EXPLICIT_CUBE
    std::string packed, geo;
    for (int d = 0; d < h.ndim; d++)
      packed.append((const char *)h.coords[d].data(), h.coords[d].size() * 8);
    f_msg(geo, 2, packed);
    f_msg(dom, 4, geo);
  }
  std::string dataset;
  f_varint(dataset, 1, h.dtype == MGB_F64 ? 1 : 0);
  f_varint(dataset, 2, 1);
  std::string err; // mode ABSOLUTE = 0
  if (!(std::isinf(h.s) && h.s > 0)) {
    f_varint(err, 2, 1); // S_NORM
    f_double(err, 3, h.s);
  }
  f_double(err, 5, h.tol);
  std::string quant;
  f_varint(quant, 1, 1); // COEFFICIENTWISE_LINEAR; PER_COEFFICIENT = 0
  f_varint(quant, 3, 3); // INT64_T
  std::string enc;
  f_varint(enc, 1, 1); // SHUFFLE
  f_varint(enc, 2, (uint64_t)h.cpu_compressor); // CPU_HUFFMAN_ZLIB = 1 / CPU_HUFFMAN_ZSTD = 2
  std::string hdr;
  f_msg(hdr, 2, ver);
  f_msg(hdr, 3, fver);
  f_msg(hdr, 4, dom);
  f_msg(hdr, 5, dataset);
  f_msg(hdr, 6, err);
  f_msg(hdr, 8, std::string()); // MULTILEVEL_COEFFICIENTS, POWER_OF_TWO_PLUS_ONE
  f_msg(hdr, 9, quant);
  f_msg(hdr, 11, enc);
  f_msg(hdr, 12, std::string()); // Device::CPU
  return with_preamble(hdr, true); // format.tpp:27-41: big-endian
}
} // namespace

std::vector<uint8_t> mgb_encode_stream_header(const mgb_header &h) {
  if (h.convention == 1)
    return encode_cpu_header(h);
  // VersionNumber: Metadata.cpp:267-271 overwrites mgard_version with the
  // FILE format version (1.0.0, CMakeLists.txt:17-19) and leaves
  // file_format_version empty; reproduced for byte-exactness.
  std::string ver;
  f_varint(ver, 1, 1);
  f_varint(ver, 2, 0);
  f_varint(ver, 3, 0);
  std::string topo;
  f_varint(topo, 1, (uint64_t)h.ndim);
  {
    std::string packed;
    for (int d = 0; d < h.ndim; d++)
      put_varint(packed, h.shape[d]);
    f_msg(topo, 2, packed);
  }
  std::string dom;
  f_msg(dom, 2, topo);
  if (!h.coords.empty()) {
    f_varint(dom, 3, 1); // EXPLICIT_CUBE
    std::string packed;
    for (int d = 0; d < h.ndim; d++)
      packed.append((const char *)h.coords[d].data(), h.coords[d].size() * 8);
    std::string geo;
    f_msg(geo, 2, packed);
    f_msg(dom, 4, geo);
  }
  std::string dataset;
  f_varint(dataset, 1, h.dtype == MGB_F64 ? 1 : 0);
  f_varint(dataset, 2, 1);
  std::string err;
  if (h.ebtype == MGB_REL)
    f_varint(err, 1, 1);
  const bool linf = std::isinf(h.s) && h.s > 0;
  if (!linf)
    f_varint(err, 2, 1);
  f_double(err, 3, h.s);
  if (h.ebtype == MGB_REL)
    f_double(err, 4, h.norm);
  f_double(err, 5, h.tol);
  std::string dd;
  f_varint(dd, 1, h.decomposed ? h.dd_method : 0); // NOOP_METHOD / MAX_DIMENSION / BLOCK / VARIABLE
  f_varint(dd, 2, h.dd_dim);
  f_varint(dd, 3, h.dd_size);
  std::string fd;
  f_varint(fd, 2, h.decomposition == 1 ? 2 : 1); // MULTIDIMENSION / ONE_DIM_AT_A_TIME _WITH_GHOST_NODES (Metadata.cpp:360-370)
  std::string quant;
  f_varint(quant, 1, 1); // COEFFICIENTWISE_LINEAR
  f_varint(quant, 3, 3); // INT64_T
  std::string enc;
  f_varint(enc, 1, h.reorder ? 1 : 0); // SHUFFLE when Config::reorder (Metadata.cpp:408-412)
  f_varint(enc, 2, h.lossless == 2 ? 5 : 3); // X_HUFFMAN / X_HUFFMAN_ZSTD (mgard.proto:139-145)
  f_varint(enc, 3, (uint64_t)h.dict_size);
  f_varint(enc, 4, (uint64_t)h.block_size);
  std::string dev;
  f_varint(dev, 1, 3); // X_CUDA
  std::string hdr;
  f_msg(hdr, 2, ver);
  f_msg(hdr, 3, std::string());
  f_msg(hdr, 4, dom);
  f_msg(hdr, 5, dataset);
  f_msg(hdr, 6, err);
  f_msg(hdr, 7, dd);
  f_msg(hdr, 8, fd);
  f_msg(hdr, 9, quant);
  f_msg(hdr, 10, std::string());
  f_msg(hdr, 11, enc);
  f_msg(hdr, 12, dev);

  return with_preamble(hdr, false);
}

uint64_t mgb_preamble_header_size(const uint8_t *pre17, size_t stream_size) {
  if (stream_size < 17 || memcmp(pre17, "MGARD", 5) != 0)
    return UINT64_MAX;
  uint64_t le = 0, be = 0;
  for (int i = 0; i < 8; i++) {
    le |= (uint64_t)pre17[5 + i] << (8 * i);
    be |= (uint64_t)pre17[5 + i] << (8 * (7 - i));
  }
  // at most one reading of a non-empty header fits the stream (the other one is a
  // byte-reversed, astronomically large number)
  const uint64_t lim = stream_size - 17;
  if (le <= lim && be <= lim)
    return le > be ? le : be;
  if (le <= lim)
    return le;
  if (be <= lim)
    return be;
  return UINT64_MAX;
}

int mgb_parse_stream_header(const uint8_t *data, size_t size, mgb_header &h,
                            uint64_t &total_bytes) {
  if (size < 17 || memcmp(data, "MGARD", 5) != 0)
    return MGB_BAD_STREAM;
  // byte order of the preamble: little-endian (MGARD-X) or big-endian (MGARD-CPU);
  // the reading whose size fits and whose CRC matches is taken and must agree with
  // the convention the header body then declares
  uint64_t hs = 0;
  bool big_endian = false, found = false;
  const uint8_t *hp = data + 17;
  for (int be = 0; be < 2 && !found; be++) {
    uint64_t cand = 0;
    uint32_t crc = 0;
    for (int i = 0; i < 8; i++)
      cand |= (uint64_t)data[5 + i] << (8 * (be ? 7 - i : i));
    for (int i = 0; i < 4; i++)
      crc |= (uint32_t)data[13 + i] << (8 * (be ? 3 - i : i));
    if (cand <= size - 17 && crc32_bytes(hp, cand) == crc) {
      hs = cand;
      big_endian = be != 0;
      found = true;
    }
  }
  if (!found)
    return MGB_BAD_STREAM;
  total_bytes = 17 + hs;

  h = mgb_header();
  h.s = 0;
  int hierarchy = 0, compressor = 0, preprocessor = 0, geometry = 0, quant_type = 0;
  uint64_t major = 0;
  std::vector<double> flat_coords;
  Reader r{hp, hp + hs};
  int f, wt;
  uint64_t v;
  Reader sub{nullptr, nullptr};
  while (r.next(f, wt, v, sub)) {
    if (wt != 2)
      continue;
    Reader m = sub, s2{nullptr, nullptr}, s3{nullptr, nullptr};
    int f2, w2;
    uint64_t v2;
    switch (f) {
    case 2:
      while (m.next(f2, w2, v2, s2))
        if (f2 == 1 && w2 == 0)
          major = v2;
      break;
    case 4: // Domain
      while (m.next(f2, w2, v2, s2)) {
        if (f2 == 2 && w2 == 2) {
          Reader t = s2;
          int f3, w3;
          uint64_t v3;
          while (t.next(f3, w3, v3, s3)) {
            if (f3 == 1 && w3 == 0)
              h.ndim = (int)v3;
            if (f3 == 2 && w3 == 2) {
              Reader pk = s3;
              int k = 0;
              while (pk.p < pk.end && pk.ok && k < MGB_MAX_DIMS)
                h.shape[k++] = pk.varint();
              if (pk.p < pk.end)
                return MGB_TOO_MANY_DIMS;
            }
          }
        } else if (f2 == 3 && w2 == 0) {
          geometry = (int)v2;
        } else if (f2 == 4 && w2 == 2) {
          Reader t = s2;
          int f3, w3;
          uint64_t v3;
          while (t.next(f3, w3, v3, s3)) {
            if (f3 == 2 && w3 == 2) {
              size_t cnt = (s3.end - s3.p) / 8;
              flat_coords.resize(cnt);
              memcpy(flat_coords.data(), s3.p, cnt * 8);
            } else if (f3 == 2 && w3 == 1) {
              flat_coords.push_back(as_double(v3));
            }
          }
        }
      }
      break;
    case 5:
      while (m.next(f2, w2, v2, s2))
        if (f2 == 1 && w2 == 0)
          h.dtype = v2 == 1 ? MGB_F64 : MGB_F32;
      break;
    case 6: {
      int mode = 0, norm = 0;
      while (m.next(f2, w2, v2, s2)) {
        if (f2 == 1 && w2 == 0) mode = (int)v2;
        if (f2 == 2 && w2 == 0) norm = (int)v2;
        if (f2 == 3 && w2 == 1) h.s = as_double(v2);
        if (f2 == 4 && w2 == 1) h.norm = as_double(v2);
        if (f2 == 5 && w2 == 1) h.tol = as_double(v2);
      }
      h.ebtype = mode == 1 ? MGB_REL : MGB_ABS;
      // Metadata.cpp:593-604: L_INFINITY => s = inf regardless of the field
      if (norm == 0)
        h.s = INFINITY;
      break;
    }
    case 7: {
      int method = 0;
      while (m.next(f2, w2, v2, s2)) {
        if (f2 == 1 && w2 == 0) method = (int)v2;
        if (f2 == 2 && w2 == 0) h.dd_dim = v2;
        if (f2 == 3 && w2 == 0) h.dd_size = v2;
      }
      if (method < 0 || method > 3)
        return MGB_BAD_STREAM;
      h.decomposed = method != 0;
      h.dd_method = method ? method : 1;
      break;
    }
    case 8:
      while (m.next(f2, w2, v2, s2))
        if (f2 == 2 && w2 == 0)
          hierarchy = (int)v2;
      break;
    case 9:
      while (m.next(f2, w2, v2, s2))
        if (f2 == 3 && w2 == 0)
          quant_type = (int)v2;
      break;
    case 11:
      while (m.next(f2, w2, v2, s2)) {
        if (f2 == 1 && w2 == 0) preprocessor = (int)v2;
        if (f2 == 2 && w2 == 0) compressor = (int)v2;
        if (f2 == 3 && w2 == 0) h.dict_size = (int)v2;
        if (f2 == 4 && w2 == 0) h.block_size = (int)v2;
      }
      break;
    default: break;
    }
    if (!m.ok)
      return MGB_BAD_STREAM;
  }
  if (!r.ok || h.ndim < 1 || h.ndim > MGB_MAX_DIMS)
    return MGB_BAD_STREAM;
  // Metadata.cpp:501-514: only the major version is checked
  if (major > 1)
    return MGB_BAD_STREAM;
  // MGARD-X multi-dimensional Huffman (+ Zstd) streams, or MGARD-CPU streams
  // (zlib or CPU Huffman + zstd payload)
  if (hierarchy == 0 && (compressor == 1 || compressor == 2)) {
    h.convention = 1;
    h.cpu_compressor = compressor;
    if (quant_type != 3 || h.ebtype != MGB_ABS)
      return MGB_BAD_STREAM;
  } else if ((hierarchy != 1 && hierarchy != 2) || (compressor != 3 && compressor != 5) || preprocessor > 1) {
    return MGB_BAD_STREAM;
  } else {
    h.reorder = preprocessor; // Metadata.cpp:693-698
    h.decomposition = hierarchy == 2 ? 1 : 0; // Metadata.cpp:620-631
  }
  // each reference reader accepts its own byte order only
  if (big_endian != (h.convention == 1))
    return MGB_BAD_STREAM;
  h.lossless = compressor == 5 ? 2 : 0;
  if (geometry == 1) {
    uint64_t tot = 0;
    for (int d = 0; d < h.ndim; d++)
      tot += h.shape[d];
    if (flat_coords.size() != tot)
      return MGB_BAD_STREAM;
    h.coords.resize(h.ndim);
    size_t off = 0;
    for (int d = 0; d < h.ndim; d++) {
      h.coords[d].assign(flat_coords.begin() + off,
                         flat_coords.begin() + off + h.shape[d]);
      off += h.shape[d];
    }
  }
  return MGB_SUCCESS;
}
