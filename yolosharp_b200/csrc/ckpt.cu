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
// The Ultralytics `.pt` pickle (ModelLoader/PickleLoader.cs) is NOT read: it needs a zip + pickle-opcode interpreter
// and the reference itself only uses it in its offline converter (Tools.TransModelFromPickle).
#include <cstdio>
#include <cstdlib>
#include <cstring>
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
    case 0: case 1: return 1;
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
  else if (ends_with(p, ".pt") || ends_with(p, ".pth")) { set_error("yb_ckpt_open: Ultralytics .pt pickles are not read natively - convert to .bin / .safetensors"); rc = YB_ERR_NOT_IMPLEMENTED; }
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
  for (const CkptTensor& t : c->tensors) {
    if (t.dtype != YB_F16 && t.dtype != YB_F32 && t.dtype != YB_BF16) continue;  // num_batches_tracked (int64) etc.
    rc = yb_load_tensor(e, t.name.c_str(), t.dtype, (int32_t)t.shape.size(), t.shape.data(), t.nbytes ? c->blob.data() + t.offset : nullptr);
    if (rc) break;
    have.push_back(t.name);
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
