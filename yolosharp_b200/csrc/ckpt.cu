// Native checkpoint ingest behind the C ABI (host code only; no CUDA calls).
//
// Replaces, for an engine that is driven without TorchSharp:
//   * `Lib.LoadModel(path)`            Utils/Lib.cs:9-54          TorchSharp `.bin`: LEB128 tensor count, then per tensor
//                                                                 {.NET BinaryWriter string (7-bit-encoded length +
//                                                                 UTF-8), LEB128 torch ScalarType, LEB128 ndim, LEB128
//                                                                 dims, raw little-endian payload}
//   * `SafetensorsLoader.ReadTensorsInfoFromFile` + `ReadByteFromFile`   ModelLoader/SafetensorsLoader.cs:7-109
//                                                                 u64 header length, JSON header
//                                                                 {name: {dtype, shape, data_offsets}}, payload
//   * `SaveWeight(path)`               Models/YoloBaseTaskModel.cs:470-490   `.bin` writer (keys containing "one2one"
//                                                                 are skipped there; the caller chooses the tensors here)
//   * `PickleLoader.Load(path)`        ModelLoader/PickleLoader.cs:21-466   torch.save archives (`.pt` / `.pth`): a ZIP with
//                                                                 stored entries `<prefix>/data.pkl` (pickle protocol 2 - 4) and
//                                                                 `<prefix>/data/<key>` (raw storages).  A pickle-opcode
//                                                                 interpreter rebuilds the object tree (dicts, lists, tuples,
//                                                                 objects with their BUILD state, `_rebuild_tensor_v2`,
//                                                                 `_rebuild_parameter`, persistent storage ids) and every
//                                                                 tensor is named by its path, as the reference's
//                                                                 ExtractTensors does (:49-88: dict keys, list indices and
//                                                                 attribute names joined by '.'); the `_parameters` /
//                                                                 `_buffers` / `_modules` levels of pickled nn.Module objects
//                                                                 are not spelled out, so a pickled module yields its
//                                                                 state_dict() names.  (The reference's own
//                                                                 ReadTensorsInfoFromFile ends in `ExtractTensors(null)`, :46,
//                                                                 i.e. returns no tensors at this commit; it is only used by
//                                                                 the offline converter Tools.TransModelFromPickle.)
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "common.cuh"

namespace yb {

struct CkptTensor {
  std::string name;
  int dtype = 0;  // torch ScalarType code (yb_dtype values for u8 / f16 / f32 / bf16; 3 = i32, 4 = i64, 7 = f64, 1 = i8, 2 = i16)
  std::vector<int64_t> shape;
  size_t offset = 0, nbytes = 0;  // into `blob`
};

static int item_size(int dt) {
  switch (dt) {
    case 0: case 1: case 11: return 1;
    case 2: case 5: case 15: return 2;
    case 3: case 6: return 4;
    case 4: case 7: return 8;
    default: return 0;
  }
}

}  // namespace yb

using namespace yb;

struct yb_ckpt {
  std::vector<uint8_t> blob;  // the whole file
  std::vector<CkptTensor> tensors;
};

namespace yb {

static bool read_file(const char* path, std::vector<uint8_t>& out) {
  FILE* f = fopen(path, "rb");
  if (!f) return false;
  fseek(f, 0, SEEK_END);
  const long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  if (n < 0) { fclose(f); return false; }
  out.resize((size_t)n);
  const size_t got = n ? fread(out.data(), 1, (size_t)n, f) : 0;
  fclose(f);
  return got == (size_t)n;
}

static bool leb(const std::vector<uint8_t>& b, size_t& pos, uint64_t& v) {
  v = 0;
  for (int shift = 0; shift < 64; shift += 7) {
    if (pos >= b.size()) return false;
    const uint8_t c = b[pos++];
    v |= (uint64_t)(c & 0x7F) << shift;
    if (!(c & 0x80)) return true;
  }
  return false;
}

static int parse_bin(yb_ckpt* c, const char* path) {
  size_t pos = 0;
  uint64_t count = 0;
  if (!leb(c->blob, pos, count) || count > 1000000) { set_error(std::string(path) + ": not a TorchSharp .bin file"); return YB_ERR_INVALID_ARG; }
  for (uint64_t i = 0; i < count; i++) {
    CkptTensor t;
    uint64_t ln = 0, dt = 0, nd = 0;
    if (!leb(c->blob, pos, ln) || pos + ln > c->blob.size()) { set_error(std::string(path) + ": truncated tensor name"); return YB_ERR_INVALID_ARG; }
    t.name.assign(reinterpret_cast<const char*>(c->blob.data() + pos), (size_t)ln);
    pos += (size_t)ln;
    if (!leb(c->blob, pos, dt) || !leb(c->blob, pos, nd) || nd > 8) { set_error(std::string(path) + ": bad header of " + t.name); return YB_ERR_INVALID_ARG; }
    t.dtype = (int)dt;
    uint64_t n = 1;
    for (uint64_t k = 0; k < nd; k++) {
      uint64_t d = 0;
      if (!leb(c->blob, pos, d)) { set_error(std::string(path) + ": truncated shape of " + t.name); return YB_ERR_INVALID_ARG; }
      t.shape.push_back((int64_t)d);
      n *= d;
    }
    const int isz = item_size(t.dtype);
    if (!isz) { set_error(std::string(path) + ": unsupported scalar type " + std::to_string(t.dtype) + " for " + t.name); return YB_ERR_NOT_IMPLEMENTED; }
    t.offset = pos;
    t.nbytes = (size_t)n * isz;
    if (pos + t.nbytes > c->blob.size()) { set_error(std::string(path) + ": truncated payload of " + t.name); return YB_ERR_INVALID_ARG; }
    pos += t.nbytes;
    c->tensors.push_back(std::move(t));
  }
  if (pos != c->blob.size()) { set_error(std::string(path) + ": trailing bytes after " + std::to_string(count) + " tensors"); return YB_ERR_INVALID_ARG; }
  return YB_OK;
}

// ---- minimal JSON reader for the safetensors header: {"name": {"dtype": "F16", "shape": [..], "data_offsets": [a, b]}, ..}
struct Json {
  const char* p;
  const char* end;
  bool ok = true;
  void ws() { while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) p++; }
  bool eat(char c) { ws(); if (p < end && *p == c) { p++; return true; } return false; }
  std::string str() {
    std::string s;
    ws();
    if (p >= end || *p != '"') { ok = false; return s; }
    p++;
    while (p < end && *p != '"') {
      if (*p == '\\' && p + 1 < end) {
        p++;
        switch (*p) {
          case 'n': s += '\n'; break;
          case 't': s += '\t'; break;
          case 'u': p += 4; s += '?'; break;  // names are ASCII in practice
          default: s += *p;
        }
        p++;
      } else {
        s += *p++;
      }
    }
    if (p >= end) { ok = false; return s; }
    p++;
    return s;
  }
  long long num() {
    ws();
    char* e = nullptr;
    const long long v = strtoll(p, &e, 10);
    if (e == p) ok = false;
    p = e;
    return v;
  }
  void skip() {  // any value
    ws();
    if (p >= end) { ok = false; return; }
    if (*p == '"') { str(); return; }
    if (*p == '{' || *p == '[') {
      const char open = *p, close = open == '{' ? '}' : ']';
      p++;
      ws();
      if (eat(close)) return;
      while (ok) {
        if (open == '{') { str(); if (!eat(':')) { ok = false; return; } }
        skip();
        if (eat(',')) continue;
        if (!eat(close)) ok = false;
        return;
      }
      return;
    }
    while (p < end && *p != ',' && *p != '}' && *p != ']') p++;  // number / true / false / null
  }
};

static int st_dtype(const std::string& s) {
  if (s == "F16") return 5;
  if (s == "F32") return 6;
  if (s == "BF16") return 15;
  if (s == "F64") return 7;
  if (s == "I64") return 4;
  if (s == "I32") return 3;
  if (s == "I16") return 2;
  if (s == "I8") return 1;
  if (s == "U8") return 0;
  return -1;
}

static int parse_safetensors(yb_ckpt* c, const char* path) {
  if (c->blob.size() < 10) { set_error(std::string(path) + ": file cannot be valid safetensors: too short"); return YB_ERR_INVALID_ARG; }
  uint64_t hs = 0;
  std::memcpy(&hs, c->blob.data(), 8);
  if (hs == 0 || hs > 100000000ull || 8 + hs > c->blob.size()) { set_error(std::string(path) + ": file cannot be valid safetensors: header length wrong"); return YB_ERR_INVALID_ARG; }
  const size_t body = 8 + (size_t)hs;
  Json j{reinterpret_cast<const char*>(c->blob.data() + 8), reinterpret_cast<const char*>(c->blob.data() + body)};
  if (!j.eat('{')) { set_error(std::string(path) + ": safetensors header is not a JSON object"); return YB_ERR_INVALID_ARG; }
  if (!j.eat('}')) {
    while (j.ok) {
      CkptTensor t;
      t.name = j.str();
      if (!j.eat(':')) { j.ok = false; break; }
      bool has_offsets = false;
      std::string dtype;
      long long o0 = 0, o1 = 0;
      if (!j.eat('{')) { j.skip(); } else if (!j.eat('}')) {
        while (j.ok) {
          const std::string key = j.str();
          if (!j.eat(':')) { j.ok = false; break; }
          if (key == "dtype") dtype = j.str();
          else if (key == "shape") {
            if (!j.eat('[')) { j.ok = false; break; }
            if (!j.eat(']')) {
              do t.shape.push_back(j.num()); while (j.eat(','));
              if (!j.eat(']')) j.ok = false;
            }
          } else if (key == "data_offsets") {
            if (!j.eat('[')) { j.ok = false; break; }
            o0 = j.num();
            if (!j.eat(',')) j.ok = false;
            o1 = j.num();
            if (!j.eat(']')) j.ok = false;
            has_offsets = true;
          } else j.skip();
          if (j.eat(',')) continue;
          if (!j.eat('}')) j.ok = false;
          break;
        }
      }
      if (j.ok && has_offsets) {  // entries without data_offsets (__metadata__) are skipped, as in the reference (:47-51)
        t.dtype = st_dtype(dtype);
        if (t.dtype < 0) { set_error(std::string(path) + ": unsupported safetensors dtype " + dtype + " for " + t.name); return YB_ERR_NOT_IMPLEMENTED; }
        size_t n = 1;
        for (int64_t d : t.shape) n *= (size_t)d;
        if (o0 < 0 || o1 < o0 || body + (size_t)o1 > c->blob.size() || (size_t)(o1 - o0) != n * item_size(t.dtype)) {
          set_error(std::string(path) + ": bad data_offsets of " + t.name);
          return YB_ERR_INVALID_ARG;
        }
        t.offset = body + (size_t)o0;
        t.nbytes = (size_t)(o1 - o0);
        c->tensors.push_back(std::move(t));
      }
      if (j.eat(',')) continue;
      if (!j.eat('}')) j.ok = false;
      break;
    }
  }
  if (!j.ok) { set_error(std::string(path) + ": malformed safetensors header"); return YB_ERR_INVALID_ARG; }
  return YB_OK;
}

static bool ends_with(const std::string& s, const char* sfx) {
  const size_t n = strlen(sfx);
  return s.size() >= n && s.compare(s.size() - n, n, sfx) == 0;
}

// ---- torch.save archives: ZIP (stored entries) + pickle -------------------------------------------------------------------
struct ZipEntry { std::string name; uint64_t data_off = 0, size = 0; int method = 0; };

static uint32_t rd32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
static uint16_t rd16(const uint8_t* p) { return (uint16_t)(p[0] | (p[1] << 8)); }
static uint64_t rd64(const uint8_t* p) { return (uint64_t)rd32(p) | ((uint64_t)rd32(p + 4) << 32); }

static bool zip_entries(const std::vector<uint8_t>& b, std::vector<ZipEntry>& out, std::string& err) {
  const size_t n = b.size();
  if (n < 22) { err = "too small for a zip archive"; return false; }
  size_t eocd = std::string::npos;
  for (size_t i = n - 22;; i--) {  // end-of-central-directory record, searched from the back (the comment is at most 64 KiB)
    if (rd32(&b[i]) == 0x06054b50u) { eocd = i; break; }
    if (i == 0 || n - i > 22 + 65535 + 1) break;
  }
  if (eocd == std::string::npos) { err = "no end-of-central-directory record"; return false; }
  uint64_t count = rd16(&b[eocd + 10]), cd_size = rd32(&b[eocd + 12]), cd_off = rd32(&b[eocd + 16]);
  if (eocd >= 20 && rd32(&b[eocd - 20]) == 0x07064b50u) {  // zip64 locator -> zip64 end-of-central-directory record
    const uint64_t z = rd64(&b[eocd - 20 + 8]);
    if (z + 56 <= n && rd32(&b[z]) == 0x06064b50u) { count = rd64(&b[z + 32]); cd_size = rd64(&b[z + 40]); cd_off = rd64(&b[z + 48]); }
  }
  if (cd_off > n || cd_size > n || cd_off + cd_size > n) { err = "central directory out of range"; return false; }
  size_t p = (size_t)cd_off;
  for (uint64_t i = 0; i < count; i++) {
    if (p + 46 > n || rd32(&b[p]) != 0x02014b50u) { err = "bad central directory entry"; return false; }
    ZipEntry e;
    e.method = rd16(&b[p + 10]);
    uint64_t csize = rd32(&b[p + 20]), usize = rd32(&b[p + 24]), lho = rd32(&b[p + 42]);
    const size_t nl = rd16(&b[p + 28]), xl = rd16(&b[p + 30]), cl = rd16(&b[p + 32]);
    if (p + 46 + nl + xl + cl > n) { err = "truncated central directory"; return false; }
    e.name.assign(reinterpret_cast<const char*>(&b[p + 46]), nl);
    for (size_t x = p + 46 + nl; x + 4 <= p + 46 + nl + xl;) {  // zip64 extended information
      const uint16_t id = rd16(&b[x]), sz = rd16(&b[x + 2]);
      if (id == 0x0001) {
        size_t q = x + 4;
        const size_t lim = std::min(x + 4 + (size_t)sz, p + 46 + nl + xl);
        if (usize == 0xffffffffu && q + 8 <= lim) { usize = rd64(&b[q]); q += 8; }
        if (csize == 0xffffffffu && q + 8 <= lim) { csize = rd64(&b[q]); q += 8; }
        if (lho == 0xffffffffu && q + 8 <= lim) { lho = rd64(&b[q]); q += 8; }
      }
      x += 4 + (size_t)sz;
    }
    if (lho > n || lho + 30 > n || rd32(&b[lho]) != 0x04034b50u) { err = "bad local header of " + e.name; return false; }
    e.data_off = lho + 30 + rd16(&b[lho + 26]) + rd16(&b[lho + 28]);
    e.size = usize;
    if (e.method == 0 && (e.data_off > n || e.size > n || e.data_off + e.size > n)) { err = "entry " + e.name + " out of range"; return false; }
    (void)csize;
    out.push_back(std::move(e));
    p += 46 + nl + xl + cl;
  }
  return true;
}

// value model of the unpickler
struct PVal;
typedef std::shared_ptr<PVal> PV;
struct PVal {
  enum Kind { NONE, BOOL, INT, FLOAT, STR, GLOBAL, TUPLE, LIST, DICT, OBJECT, STORAGE, TENSOR, MARK } kind = NONE;
  long long i = 0;
  double f = 0;
  std::string s;                                  // STR / GLOBAL name / OBJECT class / STORAGE key
  std::vector<PV> items;                          // TUPLE / LIST; OBJECT: constructor args
  std::vector<std::pair<PV, PV>> dict;            // DICT in insertion order; OBJECT: attributes (string keys)
  int dtype = 6;                                  // STORAGE / TENSOR
  PV storage;                                     // TENSOR
  long long offset = 0;                           // TENSOR: storage offset in elements
  std::vector<int64_t> shape, stride;             // TENSOR
};
static PV mk(PVal::Kind k) { PV v = std::make_shared<PVal>(); v->kind = k; return v; }

static int storage_dtype(const std::string& n) {
  struct { const char* k; int d; } T[] = {{"HalfStorage", 5}, {"BFloat16Storage", 15}, {"DoubleStorage", 7}, {"FloatStorage", 6}, {"LongStorage", 4},
                                          {"IntStorage", 3}, {"ShortStorage", 2}, {"CharStorage", 1}, {"ByteStorage", 0}, {"BoolStorage", 11}};
  for (auto& t : T) if (n.find(t.k) != std::string::npos) return t.d;
  return -1;
}

static bool unpickle(const uint8_t* d, size_t n, PV& result, std::string& err) {
  std::vector<PV> st;
  std::map<long long, PV> memo;
  auto need = [&](size_t pos, size_t k) { return pos + k <= n; };
  auto pop = [&]() { PV v = st.empty() ? mk(PVal::NONE) : st.back(); if (!st.empty()) st.pop_back(); return v; };
  auto pop_to_mark = [&]() {
    std::vector<PV> items;
    while (!st.empty() && st.back()->kind != PVal::MARK) { items.insert(items.begin(), st.back()); st.pop_back(); }
    if (!st.empty()) st.pop_back();
    return items;
  };
  auto set_item = [&](const PV& target, const PV& k, const PV& v) {
    if (target->kind != PVal::DICT && target->kind != PVal::OBJECT) return;
    if (target->kind == PVal::OBJECT && target->s.find("OrderedDict") == std::string::npos && target->s.find("dict") == std::string::npos) return;
    target->dict.push_back({k, v});
  };
  size_t p = 0;
  while (p < n) {
    const uint8_t op = d[p++];
    switch (op) {
      case 0x80: p += 1; break;                                   // PROTO
      case 0x95: p += 8; break;                                   // FRAME
      case '.': result = st.empty() ? mk(PVal::NONE) : st.back(); return true;
      case '(': st.push_back(mk(PVal::MARK)); break;
      case '0': pop(); break;                                     // POP
      case '1': pop_to_mark(); break;                             // POP_MARK
      case '2': if (!st.empty()) st.push_back(st.back()); break;  // DUP
      case 'N': st.push_back(mk(PVal::NONE)); break;
      case 0x88: case 0x89: { PV v = mk(PVal::BOOL); v->i = op == 0x88; st.push_back(v); break; }
      case 'K': { if (!need(p, 1)) goto trunc; PV v = mk(PVal::INT); v->i = d[p]; p += 1; st.push_back(v); break; }
      case 'M': { if (!need(p, 2)) goto trunc; PV v = mk(PVal::INT); v->i = rd16(d + p); p += 2; st.push_back(v); break; }
      case 'J': { if (!need(p, 4)) goto trunc; PV v = mk(PVal::INT); v->i = (int32_t)rd32(d + p); p += 4; st.push_back(v); break; }
      case 0x8a: {                                                // LONG1: little-endian two's complement
        if (!need(p, 1)) goto trunc;
        const size_t k = d[p++];
        if (!need(p, k) || k > 8) { err = "LONG1 wider than 8 bytes"; return false; }
        long long x = 0;
        for (size_t i = 0; i < k; i++) x |= (long long)d[p + i] << (8 * i);
        if (k && k < 8 && (d[p + k - 1] & 0x80)) x |= -1ll << (8 * k);
        p += k;
        PV v = mk(PVal::INT); v->i = x; st.push_back(v);
        break;
      }
      case 'G': {                                                 // BINFLOAT: big-endian double
        if (!need(p, 8)) goto trunc;
        uint64_t u = 0;
        for (int i = 0; i < 8; i++) u = (u << 8) | d[p + i];
        p += 8;
        PV v = mk(PVal::FLOAT); memcpy(&v->f, &u, 8); st.push_back(v);
        break;
      }
      case 'X': case 'T': case 'B': {                             // BINUNICODE / BINSTRING / BINBYTES (4-byte length)
        if (!need(p, 4)) goto trunc;
        const size_t k = rd32(d + p); p += 4;
        if (!need(p, k)) goto trunc;
        PV v = mk(PVal::STR); v->s.assign(reinterpret_cast<const char*>(d + p), k); p += k; st.push_back(v);
        break;
      }
      case 0x8c: case 'U': case 'C': {                            // SHORT_BINUNICODE / SHORT_BINSTRING / SHORT_BINBYTES
        if (!need(p, 1)) goto trunc;
        const size_t k = d[p++];
        if (!need(p, k)) goto trunc;
        PV v = mk(PVal::STR); v->s.assign(reinterpret_cast<const char*>(d + p), k); p += k; st.push_back(v);
        break;
      }
      case 0x8d: case 0x8e: {                                     // BINUNICODE8 / BINBYTES8
        if (!need(p, 8)) goto trunc;
        const uint64_t k = rd64(d + p); p += 8;
        if (k > n || !need(p, (size_t)k)) goto trunc;
        PV v = mk(PVal::STR); v->s.assign(reinterpret_cast<const char*>(d + p), (size_t)k); p += (size_t)k; st.push_back(v);
        break;
      }
      case 'c': {                                                 // GLOBAL: "module\nname\n"
        std::string a, b2;
        while (p < n && d[p] != '\n') a.push_back((char)d[p++]);
        p++;
        while (p < n && d[p] != '\n') b2.push_back((char)d[p++]);
        p++;
        PV v = mk(PVal::GLOBAL); v->s = a + "." + b2; st.push_back(v);
        break;
      }
      case 0x93: { PV nm = pop(), md = pop(); PV v = mk(PVal::GLOBAL); v->s = md->s + "." + nm->s; st.push_back(v); break; }  // STACK_GLOBAL
      case 'q': { if (!need(p, 1)) goto trunc; if (!st.empty()) memo[d[p]] = st.back(); p += 1; break; }
      case 'r': { if (!need(p, 4)) goto trunc; if (!st.empty()) memo[rd32(d + p)] = st.back(); p += 4; break; }
      case 0x94: { if (!st.empty()) { const long long k = (long long)memo.size(); memo[k] = st.back(); } break; }                // MEMOIZE
      case 'h': { if (!need(p, 1)) goto trunc; auto it = memo.find(d[p]); st.push_back(it == memo.end() ? mk(PVal::NONE) : it->second); p += 1; break; }
      case 'j': { if (!need(p, 4)) goto trunc; auto it = memo.find(rd32(d + p)); st.push_back(it == memo.end() ? mk(PVal::NONE) : it->second); p += 4; break; }
      case ')': st.push_back(mk(PVal::TUPLE)); break;
      case ']': st.push_back(mk(PVal::LIST)); break;
      case '}': st.push_back(mk(PVal::DICT)); break;
      case 0x8f: st.push_back(mk(PVal::LIST)); break;            // EMPTY_SET (kept as a list)
      case 't': { PV v = mk(PVal::TUPLE); v->items = pop_to_mark(); st.push_back(v); break; }
      case 'l': { PV v = mk(PVal::LIST); v->items = pop_to_mark(); st.push_back(v); break; }
      case 'd': {
        std::vector<PV> it = pop_to_mark();
        PV v = mk(PVal::DICT);
        for (size_t i = 0; i + 1 < it.size(); i += 2) v->dict.push_back({it[i], it[i + 1]});
        st.push_back(v);
        break;
      }
      case 0x85: { PV a = pop(); PV v = mk(PVal::TUPLE); v->items = {a}; st.push_back(v); break; }
      case 0x86: { PV b2 = pop(), a = pop(); PV v = mk(PVal::TUPLE); v->items = {a, b2}; st.push_back(v); break; }
      case 0x87: { PV c2 = pop(), b2 = pop(), a = pop(); PV v = mk(PVal::TUPLE); v->items = {a, b2, c2}; st.push_back(v); break; }
      case 'a': { PV x = pop(); if (!st.empty() && st.back()->kind == PVal::LIST) st.back()->items.push_back(x); break; }
      case 'e': case 0x90: {                                      // APPENDS / ADDITEMS
        std::vector<PV> it = pop_to_mark();
        if (!st.empty() && st.back()->kind == PVal::LIST) for (auto& x : it) st.back()->items.push_back(x);
        break;
      }
      case 's': { PV v = pop(), k = pop(); if (!st.empty()) set_item(st.back(), k, v); break; }
      case 'u': {
        std::vector<PV> it = pop_to_mark();
        if (!st.empty()) for (size_t i = 0; i + 1 < it.size(); i += 2) set_item(st.back(), it[i], it[i + 1]);
        break;
      }
      case 'Q': {                                                 // BINPERSID: ('storage', storage_type, key, location, numel)
        PV pid = pop();
        PV v = mk(PVal::NONE);
        if (pid->kind == PVal::TUPLE && pid->items.size() >= 5 && pid->items[0]->kind == PVal::STR && pid->items[0]->s == "storage") {
          const int dt = storage_dtype(pid->items[1]->s);
          if (dt < 0) { err = "unsupported storage type " + pid->items[1]->s; return false; }
          v = mk(PVal::STORAGE);
          v->dtype = dt;
          v->s = pid->items[2]->kind == PVal::INT ? std::to_string(pid->items[2]->i) : pid->items[2]->s;
          v->i = pid->items[4]->i;
        }
        st.push_back(v);
        break;
      }
      case 'R': case 0x81: case 0x92: {                           // REDUCE / NEWOBJ / NEWOBJ_EX
        PV kwargs = op == 0x92 ? pop() : nullptr;
        PV args = pop(), fn = pop();
        (void)kwargs;
        PV v = mk(PVal::OBJECT);
        v->s = fn->kind == PVal::GLOBAL ? fn->s : std::string();
        if (args->kind == PVal::TUPLE) v->items = args->items;
        if (op == 'R' && v->s.find("_rebuild_tensor") != std::string::npos && v->items.size() >= 4 && v->items[0]->kind == PVal::STORAGE) {
          PV t = mk(PVal::TENSOR);
          t->storage = v->items[0];
          t->dtype = t->storage->dtype;
          t->offset = v->items[1]->i;
          for (auto& x : v->items[2]->items) t->shape.push_back(x->i);
          for (auto& x : v->items[3]->items) t->stride.push_back(x->i);
          v = t;
        } else if (op == 'R' && v->s.find("_rebuild_parameter") != std::string::npos && !v->items.empty() && v->items[0]->kind == PVal::TENSOR) {
          v = v->items[0];  // Parameter(data, requires_grad, backward_hooks): the tensor itself
        }
        st.push_back(v);
        break;
      }
      case 'b': {                                                 // BUILD: obj.__setstate__(state) / obj.__dict__.update(state)
        PV state = pop();
        if (!st.empty() && st.back()->kind == PVal::OBJECT) {
          PV obj = st.back();
          const bool is_dict = obj->s.find("OrderedDict") != std::string::npos;  // a state_dict's `_metadata`: not an entry
          PV sd = state;
          if (state->kind == PVal::TUPLE && state->items.size() == 2 && state->items[0]->kind == PVal::DICT) sd = state->items[0];
          if (!is_dict && sd->kind == PVal::DICT) for (auto& kv : sd->dict) obj->dict.push_back(kv);
        }
        break;
      }
      default:
        err = "unsupported pickle opcode 0x" + std::string(1, "0123456789abcdef"[op >> 4]) + std::string(1, "0123456789abcdef"[op & 15]);
        return false;
    }
  }
trunc:
  err = "truncated pickle stream";
  return false;
}

static int pt_collect(yb_ckpt* c, const PV& v, const std::string& prefix, const std::map<std::string, const ZipEntry*>& data, const char* path,
                      int depth) {
  // a corrupted memo reference can make the object graph cyclic or heavily shared: bound the depth and the total work
  static thread_local long long budget = 0;
  if (depth == 0) budget = 4000000;
  if (!v || depth > 64) return YB_OK;
  if (--budget < 0) { set_error(std::string(path) + ": object graph too large (cyclic pickle?)"); return YB_ERR_INVALID_ARG; }
  auto join = [&](const std::string& k) { return prefix.empty() ? k : prefix + "." + k; };
  auto key_str = [](const PV& k) { return k->kind == PVal::INT ? std::to_string(k->i) : (k->kind == PVal::STR ? k->s : std::string()); };
  switch (v->kind) {
    case PVal::TENSOR: {
      if (prefix.empty()) return YB_OK;
      CkptTensor t;
      t.name = prefix;
      t.dtype = v->dtype;
      t.shape = v->shape;
      const int isz = item_size(t.dtype);
      if (!isz || !v->storage || v->stride.size() != t.shape.size() || t.shape.size() > 16 || v->offset < 0) {
        set_error(std::string(path) + ": malformed tensor record for " + t.name);
        return YB_ERR_INVALID_ARG;
      }
      uint64_t numel = 1;
      for (auto dsz : t.shape) {
        if (dsz < 0 || (dsz > 0 && numel > (uint64_t)1 << 40)) { set_error(std::string(path) + ": bad shape of " + t.name); return YB_ERR_INVALID_ARG; }
        numel *= (uint64_t)dsz;
      }
      if (numel > (uint64_t)1 << 40 || (uint64_t)v->offset > (uint64_t)1 << 40) { set_error(std::string(path) + ": bad extent of " + t.name); return YB_ERR_INVALID_ARG; }
      int64_t expect = 1;  // contiguous row-major strides (dimensions of extent 1 may carry any stride)
      for (int k = (int)t.shape.size() - 1; k >= 0; k--) {
        if (t.shape[k] != 1 && numel && v->stride[k] != expect) { set_error(std::string(path) + ": " + t.name + " is not contiguous"); return YB_ERR_NOT_IMPLEMENTED; }
        expect *= t.shape[k];
      }
      auto it = data.find(v->storage->s);
      if (it == data.end()) { set_error(std::string(path) + ": storage '" + v->storage->s + "' of " + t.name + " is not in the archive"); return YB_ERR_INVALID_ARG; }
      if (it->second->method != 0) { set_error(std::string(path) + ": compressed zip entries are not supported (torch.save stores)"); return YB_ERR_NOT_IMPLEMENTED; }
      t.nbytes = (size_t)numel * isz;
      const uint64_t off = (uint64_t)v->offset * isz;
      if (off + t.nbytes > it->second->size) { set_error(std::string(path) + ": " + t.name + " exceeds its storage"); return YB_ERR_INVALID_ARG; }
      t.offset = (size_t)(it->second->data_off + off);
      c->tensors.push_back(std::move(t));
      return YB_OK;
    }
    case PVal::TUPLE: case PVal::LIST:
      for (size_t i = 0; i < v->items.size(); i++)
        if (int rc = pt_collect(c, v->items[i], join(std::to_string(i)), data, path, depth + 1)) return rc;
      return YB_OK;
    case PVal::DICT:
      for (auto& kv : v->dict) {
        const std::string k = key_str(kv.first);
        if (k.empty()) continue;
        if (int rc = pt_collect(c, kv.second, join(k), data, path, depth + 1)) return rc;
      }
      return YB_OK;
    case PVal::OBJECT:
      for (auto& kv : v->dict) {
        const std::string k = key_str(kv.first);
        if (k.empty()) continue;
        // nn.Module internals: parameters, buffers and sub-modules are named as state_dict() names them
        const bool transparent = k == "_parameters" || k == "_buffers" || k == "_modules";
        if (int rc = pt_collect(c, kv.second, transparent ? prefix : join(k), data, path, depth + 1)) return rc;
      }
      return YB_OK;
    default:
      return YB_OK;
  }
}

static int parse_pt(yb_ckpt* c, const char* path) {
  std::vector<ZipEntry> entries;
  std::string err;
  if (!zip_entries(c->blob, entries, err)) { set_error(std::string(path) + ": " + err); return YB_ERR_INVALID_ARG; }
  const ZipEntry* pkl = nullptr;
  for (auto& e : entries) if (e.name == "data.pkl" || ends_with(e.name, "/data.pkl")) { pkl = &e; break; }
  if (!pkl) { set_error(std::string(path) + ": no data.pkl in the archive"); return YB_ERR_INVALID_ARG; }  // reference: ArgumentException (:34)
  if (pkl->method != 0) { set_error(std::string(path) + ": data.pkl is compressed"); return YB_ERR_NOT_IMPLEMENTED; }
  if (pkl->size < 2 || c->blob[pkl->data_off] != 0x80) { set_error(std::string(path) + ": not a valid pickle"); return YB_ERR_INVALID_ARG; }  // (:45)
  const std::string prefix = pkl->name.substr(0, pkl->name.size() - 8);  // "archive/" or ""
  std::map<std::string, const ZipEntry*> data;
  for (auto& e : entries)
    if (e.name.compare(0, prefix.size() + 5, prefix + "data/") == 0) data[e.name.substr(prefix.size() + 5)] = &e;
  PV root;
  if (!unpickle(c->blob.data() + pkl->data_off, (size_t)pkl->size, root, err)) { set_error(std::string(path) + ": " + err); return YB_ERR_INVALID_ARG; }
  return pt_collect(c, root, "", data, path, 0);
}

static void put_leb(FILE* f, uint64_t v) {
  do {
    uint8_t b = v & 0x7F;
    v >>= 7;
    if (v) b |= 0x80;
    fputc(b, f);
  } while (v);
}

}  // namespace yb

extern "C" {

int32_t yb_ckpt_open(const char* path, yb_ckpt** out) {
  if (!path || !out) { set_error("yb_ckpt_open: null argument"); return YB_ERR_INVALID_ARG; }
  *out = nullptr;
  yb_ckpt* c = new yb_ckpt();
  if (!read_file(path, c->blob)) {
    set_error(std::string("yb_ckpt_open: cannot read ") + path);  // reference: FileNotFoundException
    delete c;
    return YB_ERR_INVALID_ARG;
  }
  const std::string p(path);
  int rc;
  if (ends_with(p, ".safetensors")) rc = parse_safetensors(c, path);
  else if (ends_with(p, ".pt") || ends_with(p, ".pth")) rc = parse_pt(c, path);
  else rc = parse_bin(c, path);
  if (rc) { delete c; return rc; }
  *out = c;
  return YB_OK;
}

int32_t yb_ckpt_count(const yb_ckpt* c) { return c ? (int32_t)c->tensors.size() : 0; }

int32_t yb_ckpt_tensor(const yb_ckpt* c, int32_t i, const char** name, int32_t* dtype, int32_t* ndim, const int64_t** shape,
                       const void** data, int64_t* nbytes) {
  if (!c || i < 0 || i >= (int32_t)c->tensors.size()) { set_error("yb_ckpt_tensor: bad index"); return YB_ERR_INVALID_ARG; }
  const CkptTensor& t = c->tensors[i];
  if (name) *name = t.name.c_str();
  if (dtype) *dtype = t.dtype;
  if (ndim) *ndim = (int32_t)t.shape.size();
  if (shape) *shape = t.shape.data();
  if (data) *data = c->blob.data() + t.offset;
  if (nbytes) *nbytes = (int64_t)t.nbytes;
  return YB_OK;
}

void yb_ckpt_close(yb_ckpt* c) { delete c; }

int32_t yb_load_checkpoint(yb_engine* e, const char* path, int32_t* n_loaded, int32_t* n_missing) {
  if (!e || !path) { set_error("yb_load_checkpoint: null argument"); return YB_ERR_INVALID_ARG; }
  yb_ckpt* c = nullptr;
  int rc = yb_ckpt_open(path, &c);
  if (rc) return rc;
  int loaded = 0;
  std::vector<std::string> have;
  // an Ultralytics checkpoint pickles {'model': <model object>}: its tensors arrive as "model.<state_dict key>".  When no
  // tensor of the file carries an expected name as it stands but the names do after dropping that first "model." level,
  // drop it (the reference goes the other way round with PickleLoader.Load(fileName, addString), PickleLoader.cs:438-465)
  std::vector<std::string> expected;
  for (int i = 0, n = yb_num_expected_tensors(e); i < n; i++) expected.push_back(yb_expected_tensor_name(e, i));
  auto is_expected = [&](const std::string& k) { return std::find(expected.begin(), expected.end(), k) != expected.end(); };
  int direct = 0, stripped = 0;
  for (const CkptTensor& t : c->tensors) {
    if (is_expected(t.name)) direct++;
    else if (t.name.compare(0, 6, "model.") == 0 && is_expected(t.name.substr(6))) stripped++;
  }
  const bool strip = direct == 0 && stripped > 0;
  for (const CkptTensor& t : c->tensors) {
    if (t.dtype != YB_F16 && t.dtype != YB_F32 && t.dtype != YB_BF16) continue;  // num_batches_tracked (int64) etc.
    const std::string name = strip && t.name.compare(0, 6, "model.") == 0 ? t.name.substr(6) : t.name;
    rc = yb_load_tensor(e, name.c_str(), t.dtype, (int32_t)t.shape.size(), t.shape.data(), t.nbytes ? c->blob.data() + t.offset : nullptr);
    if (rc) break;
    have.push_back(name);
    loaded++;
  }
  int missing = 0;
  if (!rc) {
    const int n = yb_num_expected_tensors(e);
    for (int i = 0; i < n; i++) {
      const char* want = yb_expected_tensor_name(e, i);
      bool found = false;
      for (const auto& h : have)
        if (h == want) { found = true; break; }
      if (!found) missing++;
    }
  }
  yb_ckpt_close(c);
  if (n_loaded) *n_loaded = loaded;
  if (n_missing) *n_missing = missing;
  return rc;
}

int32_t yb_ckpt_write_bin(const char* path, int32_t count, const char* const* names, const int32_t* dtypes, const int32_t* ndims,
                          const int64_t* const* shapes, const void* const* datas) {
  if (!path || count < 0 || (count > 0 && (!names || !dtypes || !ndims || !shapes || !datas))) { set_error("yb_ckpt_write_bin: null argument"); return YB_ERR_INVALID_ARG; }
  FILE* f = fopen(path, "wb");
  if (!f) { set_error(std::string("yb_ckpt_write_bin: cannot create ") + path); return YB_ERR_INVALID_ARG; }
  put_leb(f, (uint64_t)count);
  for (int i = 0; i < count; i++) {
    const int isz = item_size(dtypes[i]);
    if (!isz || ndims[i] < 0 || ndims[i] > 8) { fclose(f); set_error("yb_ckpt_write_bin: unsupported dtype / rank"); return YB_ERR_INVALID_ARG; }
    const size_t ln = strlen(names[i]);
    put_leb(f, ln);  // .NET BinaryWriter.Write(string): 7-bit-encoded byte length + UTF-8
    fwrite(names[i], 1, ln, f);
    put_leb(f, (uint64_t)dtypes[i]);
    put_leb(f, (uint64_t)ndims[i]);
    size_t n = 1;
    for (int k = 0; k < ndims[i]; k++) {
      put_leb(f, (uint64_t)shapes[i][k]);
      n *= (size_t)shapes[i][k];
    }
    if (n && fwrite(datas[i], isz, n, f) != n) { fclose(f); set_error("yb_ckpt_write_bin: write failed"); return YB_ERR_INVALID_ARG; }
  }
  fclose(f);
  return YB_OK;
}

}  // extern "C"
