#include "weights.h"

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>

namespace spfe {
namespace {

bool read_file(const std::string &path, std::vector<uint8_t> &buf, std::string &err) {
  FILE *f = fopen(path.c_str(), "rb");
  if (!f) { err = "cannot open weights file '" + path + "'"; return false; }
  fseek(f, 0, SEEK_END);
  long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  buf.resize(n > 0 ? static_cast<size_t>(n) : 0);
  size_t got = buf.empty() ? 0 : fread(buf.data(), 1, buf.size(), f);
  fclose(f);
  if (got != buf.size() || buf.size() < 8) { err = "short read on '" + path + "'"; return false; }
  return true;
}

inline uint16_t rd16(const uint8_t *p) { return static_cast<uint16_t>(p[0] | (p[1] << 8)); }
inline uint32_t rd32(const uint8_t *p) { return p[0] | (p[1] << 8) | (p[2] << 16) | (static_cast<uint32_t>(p[3]) << 24); }

// Element count of a tensor read from a file: every dimension in 1 .. 2^20 and the product below 2^28 (the largest
// SuperPoint tensor has 589 824 elements), so that no later size computation can wrap.
bool checked_numel(const std::vector<int> &dims, size_t &numel) {
  if (dims.empty() || dims.size() > 4) return false;
  uint64_t n = 1;
  for (int d : dims) {
    if (d <= 0 || d > (1 << 20)) return false;
    n *= static_cast<uint64_t>(d);
    if (n > (1u << 28)) return false;
  }
  numel = static_cast<size_t>(n);
  return true;
}

// ---- .spw ------------------------------------------------------------------
bool parse_spw(const std::vector<uint8_t> &buf, WeightMap &out, std::string &err) {
  const uint32_t n = rd32(&buf[4]);
  size_t off = 8;
  for (uint32_t i = 0; i < n; i++) {
    if (off + 52 > buf.size()) { err = "spw: truncated header"; return false; }
    char name[33];
    memcpy(name, &buf[off], 32);
    name[32] = 0;
    const uint32_t ndim = rd32(&buf[off + 32]);
    if (ndim < 1 || ndim > 4) { err = "spw: bad ndim"; return false; }
    HostTensor t;
    for (uint32_t d = 0; d < ndim; d++) t.dims.push_back(static_cast<int>(rd32(&buf[off + 36 + 4 * d])));
    off += 52;
    size_t cnt = 0;
    if (!checked_numel(t.dims, cnt)) { err = "spw: bad tensor dimensions"; return false; }
    if (cnt > (buf.size() - off) / 4) { err = "spw: truncated tensor data"; return false; }
    t.data.resize(cnt);
    memcpy(t.data.data(), &buf[off], cnt * 4);
    off += cnt * 4;
    out[name] = std::move(t);
  }
  return true;
}

// ---- minimal JSON ------------------------------------------------------------
struct JVal {
  enum Kind { NUL, BOOL, NUM, STR, ARR, OBJ } kind = NUL;
  std::string s;  // STR / NUM text
  std::vector<JVal> a;
  std::vector<std::pair<std::string, JVal>> o;
  const JVal *get(const char *key) const {
    for (auto &kv : o) if (kv.first == key) return &kv.second;
    return nullptr;
  }
};

struct JParser {
  const char *p, *end;
  bool ok = true;
  int depth = 0;  // nesting of the value being parsed (a crafted file must not overflow the stack)
  void ws() { while (p < end && (*p == ' ' || *p == '\n' || *p == '\r' || *p == '\t')) p++; }
  std::string str() {
    std::string r;
    p++;  // opening quote
    while (p < end && *p != '"') {
      if (*p == '\\' && p + 1 < end) { p++; r.push_back(*p == 'n' ? '\n' : *p == 't' ? '\t' : *p); }
      else r.push_back(*p);
      p++;
    }
    if (p >= end) ok = false; else p++;
    return r;
  }
  JVal val() {
    JVal v;
    if (++depth > 64) ok = false;
    if (ok) parse(v);
    depth--;
    return v;
  }
  void parse(JVal &v) {
    ws();
    if (p >= end) { ok = false; return; }
    if (*p == '{') {
      v.kind = JVal::OBJ; p++; ws();
      if (p < end && *p == '}') { p++; return; }
      while (ok) {
        ws();
        if (p >= end || *p != '"') { ok = false; break; }
        std::string k = str();
        ws();
        if (p >= end || *p != ':') { ok = false; break; }
        p++;
        v.o.emplace_back(std::move(k), val());
        ws();
        if (p < end && *p == ',') { p++; continue; }
        if (p < end && *p == '}') { p++; break; }
        ok = false;
      }
    } else if (*p == '[') {
      v.kind = JVal::ARR; p++; ws();
      if (p < end && *p == ']') { p++; return; }
      while (ok) {
        v.a.push_back(val());
        ws();
        if (p < end && *p == ',') { p++; continue; }
        if (p < end && *p == ']') { p++; break; }
        ok = false;
      }
    } else if (*p == '"') {
      v.kind = JVal::STR; v.s = str();
    } else if (end - p >= 4 && !strncmp(p, "true", 4)) { v.kind = JVal::BOOL; v.s = "1"; p += 4; }
    else if (end - p >= 5 && !strncmp(p, "false", 5)) { v.kind = JVal::BOOL; v.s = "0"; p += 5; }
    else if (end - p >= 4 && !strncmp(p, "null", 4)) { p += 4; }
    else {
      v.kind = JVal::NUM;
      while (p < end && (strchr("+-.eE", *p) || (*p >= '0' && *p <= '9'))) v.s.push_back(*p++);
      if (v.s.empty()) ok = false;
    }
  }
};

// ---- legacy PyTorch-1.0 zip archive -------------------------------------------
struct ZipEntry { size_t data_off, size; };

bool parse_zip(const std::vector<uint8_t> &buf, std::map<std::string, ZipEntry> &entries, std::string &err) {
  // End-of-central-directory record: signature 0x06054b50, scan backwards.
  size_t eocd = std::string::npos;
  for (size_t i = buf.size() >= 22 ? buf.size() - 22 : 0;; i--) {
    if (rd32(&buf[i]) == 0x06054b50u) { eocd = i; break; }
    if (i == 0 || buf.size() - i > 66000) break;
  }
  if (eocd == std::string::npos) { err = "zip: end-of-central-directory not found"; return false; }
  if (eocd + 22 > buf.size()) { err = "zip: truncated end-of-central-directory"; return false; }
  const uint16_t n = rd16(&buf[eocd + 10]);
  size_t cd = rd32(&buf[eocd + 16]);
  for (uint16_t i = 0; i < n; i++) {
    if (cd + 46 > buf.size() || rd32(&buf[cd]) != 0x02014b50u) { err = "zip: bad central directory"; return false; }
    const uint16_t method = rd16(&buf[cd + 10]);
    const uint32_t csize = rd32(&buf[cd + 20]), usize = rd32(&buf[cd + 24]);
    const uint16_t nlen = rd16(&buf[cd + 28]), xlen = rd16(&buf[cd + 30]), clen = rd16(&buf[cd + 32]);
    const size_t lho = rd32(&buf[cd + 42]);  // all offset arithmetic in size_t: 32-bit sums could wrap
    if (cd + 46 + static_cast<size_t>(nlen) + xlen + clen > buf.size()) { err = "zip: central directory entry out of range"; return false; }
    std::string name(reinterpret_cast<const char *>(&buf[cd + 46]), nlen);
    if (method != 0 || csize != usize) { err = "zip: entry '" + name + "' is compressed (expected stored)"; return false; }
    if (lho + 30 > buf.size() || rd32(&buf[lho]) != 0x04034b50u) { err = "zip: bad local header"; return false; }
    const size_t data_off = lho + 30 + rd16(&buf[lho + 26]) + rd16(&buf[lho + 28]);
    if (data_off > buf.size() || usize > buf.size() - data_off) { err = "zip: entry out of range"; return false; }
    entries[name] = ZipEntry{data_off, usize};
    cd += 46 + static_cast<size_t>(nlen) + xlen + clen;
  }
  return true;
}

bool parse_legacy_pt(const std::vector<uint8_t> &buf, WeightMap &out, std::string &err) {
  std::map<std::string, ZipEntry> ent;
  if (!parse_zip(buf, ent, err)) return false;
  std::string root;
  for (auto &kv : ent) {
    const size_t k = kv.first.find("/model.json");
    if (k != std::string::npos && k + 11 == kv.first.size()) root = kv.first.substr(0, k);
  }
  if (root.empty()) { err = "legacy archive: model.json not found"; return false; }
  const ZipEntry mj = ent[root + "/model.json"];
  JParser jp{reinterpret_cast<const char *>(&buf[mj.data_off]), reinterpret_cast<const char *>(&buf[mj.data_off]) + mj.size};
  const JVal doc = jp.val();
  const JVal *tensors = doc.get("tensors");
  const JVal *mm = doc.get("mainModule");
  const JVal *subs = mm ? mm->get("submodules") : nullptr;
  if (!jp.ok || !tensors || !subs || tensors->kind != JVal::ARR) { err = "legacy archive: malformed model.json"; return false; }
  for (const JVal &sub : subs->a) {
    const JVal *nm = sub.get("name"), *params = sub.get("parameters");
    if (!nm || !params) continue;
    for (const JVal &par : params->a) {
      const JVal *pid = par.get("tensorId"), *pn = par.get("name");
      if (!pid || !pn) continue;
      const long long id_ll = atoll(pid->s.c_str());
      if (id_ll < 0 || static_cast<size_t>(id_ll) >= tensors->a.size()) { err = "legacy archive: tensorId out of range"; return false; }
      const size_t id = static_cast<size_t>(id_ll);
      const JVal &t = tensors->a[id];
      const JVal *dims = t.get("dims"), *dt = t.get("dataType"), *data = t.get("data"), *off = t.get("offset");
      const JVal *key = data ? data->get("key") : nullptr;
      if (!dims || !key || !dt || dt->s != "FLOAT") { err = "legacy archive: unsupported tensor record"; return false; }
      HostTensor ht;
      for (const JVal &d : dims->a) ht.dims.push_back(atoi(d.s.c_str()));
      auto it = ent.find(root + "/" + key->s);
      const long long eoff_ll = off ? atoll(off->s.c_str()) : 0;
      size_t cnt = 0;
      if (!checked_numel(ht.dims, cnt)) { err = "legacy archive: bad tensor dimensions"; return false; }
      if (it == ent.end() || eoff_ll < 0 || static_cast<unsigned long long>(eoff_ll) > it->second.size / 4 ||
          cnt > it->second.size / 4 - static_cast<size_t>(eoff_ll)) { err = "legacy archive: tensor data missing"; return false; }
      const size_t eoff = static_cast<size_t>(eoff_ll);
      ht.data.resize(cnt);
      memcpy(ht.data.data(), &buf[it->second.data_off + eoff * 4], cnt * 4);
      out[nm->s + "." + pn->s] = std::move(ht);
    }
  }
  return true;
}

}  // namespace

bool load_weights(const std::string &path, WeightMap &out, std::string &err) {
  std::vector<uint8_t> buf;
  if (!read_file(path, buf, err)) return false;
  out.clear();
  bool ok;
  if (!memcmp(buf.data(), "SPW1", 4)) ok = parse_spw(buf, out, err);
  else if (rd32(buf.data()) == 0x04034b50u) ok = parse_legacy_pt(buf, out, err);
  else { err = "unrecognised weights format (expected legacy superpoint.pt zip or SPW1)"; return false; }
  if (!ok) return false;
  // Architecture check: sp_extractor.cpp:16-43.
  static const struct { const char *name; int co, ci, k; } plan[] = {
      {"conv1a", 64, 1, 3},    {"conv1b", 64, 64, 3},   {"conv2a", 64, 64, 3},   {"conv2b", 64, 64, 3},
      {"conv3a", 128, 64, 3},  {"conv3b", 128, 128, 3}, {"conv4a", 128, 128, 3}, {"conv4b", 128, 128, 3},
      {"convPa", 256, 128, 3}, {"convPb", 65, 256, 1},  {"convDa", 256, 128, 3}, {"convDb", 256, 256, 1}};
  for (auto &l : plan) {
    auto w = out.find(std::string(l.name) + ".weight"), b = out.find(std::string(l.name) + ".bias");
    if (w == out.end() || b == out.end()) { err = std::string("weights: missing ") + l.name; return false; }
    const std::vector<int> want = {l.co, l.ci, l.k, l.k};
    if (w->second.dims != want || b->second.numel() != static_cast<size_t>(l.co)) {
      err = std::string("weights: wrong shape for ") + l.name;
      return false;
    }
  }
  return true;
}

}  // namespace spfe
