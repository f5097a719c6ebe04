// Native reader of rating files.  Same rules as FileIO.loadDataSet of the reference
// (util/io.py:31-76) with the default delimiter set and as Rating.__generateSet (data/rating.py:33-54):
//   * optional header line skipped; every other line is stripped, then split at EVERY single ' ', ','
//     or tab (re.split(' |,|\t'): two separators in a row make an empty field);
//   * columns pick user / item / rating (rating column optional: 1.0); when binarising, lines with
//     rating < threshold are dropped and the rest become 1.0;
//   * user and item names get dense ids in order of FIRST APPEARANCE.
// Anything this reader is not sure about (a short line, a rating strtod does not consume entirely)
// is reported as an error so that the caller can fall back to the Python path and fail there with the
// reference's own exception.
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <string_view>
#include <vector>

#include "common.h"

struct qrec_text_table {
  std::vector<int32_t> u, i;
  std::vector<double> r;
  std::vector<std::string> names[2];      // 0 users, 1 items, first-appearance order
};

namespace {

inline bool is_strip(unsigned char c) { return c == ' ' || (c >= '\t' && c <= '\r') || (c >= 0x1c && c <= 0x1f); }
inline bool is_sep(char c) { return c == ' ' || c == ',' || c == '\t'; }

// "12", "3.5", ".5", "4." -> value; both the digit string (< 2^53) and the power of ten (<= 10^22) are exact
// doubles, so their quotient is the correctly rounded result -- what strtod / Python's float() return
inline bool parse_short_decimal(std::string_view s, double* out) {
  static const double P10[] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11,
                               1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
  if (s.empty() || s.size() > 17) return false;
  uint64_t digits = 0;
  int n_digits = 0, frac = 0;
  bool dot = false;
  for (char c : s) {
    if (c >= '0' && c <= '9') {
      digits = digits * 10 + (uint64_t)(c - '0');
      ++n_digits;
      if (dot) ++frac;
    } else if (c == '.' && !dot) {
      dot = true;
    } else {
      return false;
    }
  }
  if (n_digits == 0 || n_digits > 15) return false;
  *out = (double)digits / P10[frac];
  return true;
}

// name -> dense id in order of first appearance: open addressing over the name list, no allocation on a hit
struct Vocab {
  std::vector<std::string>* names;
  std::vector<uint64_t> hashes;          // per id
  std::vector<int32_t> slots;            // power-of-two table of ids, -1 = empty
  explicit Vocab(std::vector<std::string>* n) : names(n), slots(1 << 16, -1) {}
  static uint64_t hash(std::string_view s) {
    uint64_t h = 0xcbf29ce484222325ull;                   // FNV-1a, then a finaliser for the low bits
    for (unsigned char c : s) h = (h ^ c) * 0x100000001b3ull;
    h ^= h >> 32;
    return h * 0x9e3779b97f4a7c15ull;
  }
  void grow() {
    std::vector<int32_t> bigger(slots.size() * 2, -1);
    const size_t mask = bigger.size() - 1;
    for (size_t id = 0; id < hashes.size(); ++id) {
      size_t k = (size_t)(hashes[id] >> 20) & mask;
      while (bigger[k] != -1) k = (k + 1) & mask;
      bigger[k] = (int32_t)id;
    }
    slots.swap(bigger);
  }
  int32_t id_of(std::string_view s) {
    const uint64_t h = hash(s);
    size_t mask = slots.size() - 1;
    size_t k = (size_t)(h >> 20) & mask;
    while (slots[k] != -1) {
      const int32_t id = slots[k];
      if (hashes[(size_t)id] == h && (*names)[(size_t)id] == s) return id;
      k = (k + 1) & mask;
    }
    const int32_t id = (int32_t)names->size();
    names->emplace_back(s);
    hashes.push_back(h);
    slots[k] = id;
    if (hashes.size() * 2 > slots.size()) grow();
    return id;
  }
};

}  // namespace

extern "C" {

qrec_text_table* qrec_text_load(const char* path, int32_t col_u, int32_t col_i, int32_t col_r, int32_t header,
                                int32_t binarize, double threshold) {
  if (!path || col_u < 0 || col_i < 0 || (binarize && col_r < 0)) {
    qrec::set_error("qrec_text_load: bad arguments");
    return nullptr;
  }
  FILE* fh = std::fopen(path, "rb");
  if (!fh) {
    qrec::set_error("qrec_text_load: cannot open %s: %s", path, std::strerror(errno));
    return nullptr;
  }
  std::string data;
  {
    long size = 0;
    if (std::fseek(fh, 0, SEEK_END) == 0 && (size = std::ftell(fh)) > 0 && std::fseek(fh, 0, SEEK_SET) == 0) {
      data.resize((size_t)size);
      const size_t got = std::fread(&data[0], 1, (size_t)size, fh);
      data.resize(got);
    } else {                                            // not seekable: read in pieces
      std::rewind(fh);
      char buf[1 << 16];
      size_t got;
      while ((got = std::fread(buf, 1, sizeof buf, fh)) > 0) data.append(buf, got);
    }
    std::fclose(fh);
  }
  auto* t = new qrec_text_table();
  Vocab users(&t->names[0]), items(&t->names[1]);
  const int need = std::max(std::max(col_u, col_i), col_r) + 1;
  std::vector<std::string_view> f;
  const char* p = data.data();
  const char* end = p + data.size();
  long long lineno = 0;
  while (p < end) {
    const char* nl = static_cast<const char*>(std::memchr(p, '\n', (size_t)(end - p)));
    const char* le = nl ? nl : end;
    const char* next = nl ? nl + 1 : end;
    const long long this_line = lineno++;
    if (header && this_line == 0) {
      p = next;
      continue;
    }
    const char* a = p;
    const char* b = le;
    while (a < b && is_strip((unsigned char)*a)) ++a;
    while (b > a && is_strip((unsigned char)b[-1])) --b;
    f.clear();
    const char* s = a;
    for (const char* c = a; c <= b; ++c) {
      if (c == b || is_sep(*c)) {
        f.emplace_back(s, (size_t)(c - s));
        s = c + 1;
      }
    }
    if ((int)f.size() < need) {
      qrec::set_error("qrec_text_load: line %lld of %s has %d fields, %d needed", this_line + 1, path, (int)f.size(), need);
      delete t;
      return nullptr;
    }
    double rating = 1.0;
    if (col_r >= 0 && parse_short_decimal(f[(size_t)col_r], &rating)) {
      // digits[.digits] with < 2^53 / <= 22 decimals: one exact division, the same double float() gives
    } else if (col_r >= 0) {
      const std::string tok(f[(size_t)col_r]);
      char* stop = nullptr;
      errno = 0;
      rating = std::strtod(tok.c_str(), &stop);
      // float() also accepts forms strtod does not (underscores) and rejects some it accepts (hex):
      // only plain decimal tokens are taken here
      bool plain = !tok.empty() && stop == tok.c_str() + tok.size();
      for (char ch : tok)
        if (!((ch >= '0' && ch <= '9') || ch == '.' || ch == '-' || ch == '+' || ch == 'e' || ch == 'E')) plain = false;
      if (!plain) {
        qrec::set_error("qrec_text_load: line %lld of %s: rating '%s' is not a plain decimal number", this_line + 1, path, tok.c_str());
        delete t;
        return nullptr;
      }
    }
    if (binarize) {
      if (rating < threshold) {
        p = next;
        continue;
      }
      rating = 1.0;
    }
    t->u.push_back(users.id_of(f[(size_t)col_u]));
    t->i.push_back(items.id_of(f[(size_t)col_i]));
    t->r.push_back(rating);
    p = next;
  }
  return t;
}

int64_t qrec_text_rows(const qrec_text_table* t) { return t ? (int64_t)t->u.size() : -1; }

int32_t qrec_text_vocab_size(const qrec_text_table* t, int32_t which) {
  return (t && (which == 0 || which == 1)) ? (int32_t)t->names[which].size() : -1;
}

int qrec_text_copy(const qrec_text_table* t, int32_t* u, int32_t* i, double* r) {
  QREC_REQUIRE(t && (t->u.empty() || (u && i && r)), "qrec_text_copy: null pointer");
  if (!t->u.empty()) {
    std::memcpy(u, t->u.data(), t->u.size() * sizeof(int32_t));
    std::memcpy(i, t->i.data(), t->i.size() * sizeof(int32_t));
    std::memcpy(r, t->r.data(), t->r.size() * sizeof(double));
  }
  return QREC_OK;
}

// names joined by '\n' (a name cannot contain one); returns the byte count, copies when buf != null
int64_t qrec_text_names(const qrec_text_table* t, int32_t which, char* buf, int64_t capacity) {
  if (!t || (which != 0 && which != 1)) return -1;
  int64_t total = 0;
  for (const auto& s : t->names[which]) total += (int64_t)s.size() + 1;
  if (total > 0) --total;                                 // no trailing separator
  if (buf) {
    if (capacity < total) return -1;
    char* o = buf;
    for (size_t k = 0; k < t->names[which].size(); ++k) {
      if (k) *o++ = '\n';
      std::memcpy(o, t->names[which][k].data(), t->names[which][k].size());
      o += t->names[which][k].size();
    }
  }
  return total;
}

void qrec_text_free(qrec_text_table* t) { delete t; }

}  // extern "C"
