// oa_jsonl.hpp — OpenAlex `works` JSON-lines -> {"id","document"} JSON-lines.
//
// Stands in for the `./oa_jsonl` stage of the reference pipeline (/root/reference/Makefile:64,
// program /root/reference/oa_jsonl.c:351-414).  Same observable behaviour on well-formed input,
// byte for byte (tests/test_oa_jsonl_cpu.py checks it against the reference binary itself,
// oracle/_ref/oa_jsonl), but built differently: the reference walks one NUL-terminated line at a
// time, fgetc by fgetc, patching terminators into the line; this works on a read-only byte range
// with (pointer, length) views, finds string ends with memchr, and converts disjoint line ranges
// of one buffer on several threads.  Plain C++17, no CUDA: shared by oa_jsonl.cu (C ABI) and
// oa_jsonl_cli.cpp (the drop-in `oa_jsonl` executable).
//
// Behaviour kept (reference line in brackets):
//   * a record is emitted only if it has an abstract_inverted_index that builds a non-empty
//     abstract [oa_jsonl.c:387-392, 402-410]; `language` present and null or != "en" drops it
//     [:378-385]; no `language` key at all keeps it;
//   * document = title + ' ' + abstract when title is a string (even an empty one), else abstract
//     alone [:402-410]; a missing id prints as "(null)" (glibc printf("%s", NULL));
//   * abstract = words placed at their positions, gaps skipped without a doubled space, one ' '
//     after every placed word except the one in the last slot [:260-282]; a position listed twice
//     keeps the word parsed last [:232-250];
//   * strings are passed through still JSON-escaped [:401]; keys compare on their raw bytes;
//   * whitespace is ' ', '\t', '\r' only [:43-47]; an empty line ends the conversion [:363-366];
//     a last line without '\n' is converted [:333-349].
// Deliberate differences: malformed records raise an error carrying the line number instead of
// tripping assert()/reading out of bounds; negative or absurd (> 2^24) word positions are errors
// (the reference writes out of bounds / reallocs gigabytes).
#pragma once

#include <cstdint>
#include <cstring>
#include <string>
#include <string_view>
#include <thread>
#include <vector>

namespace absb {
namespace oa {

struct ParseError {
  int64_t line;  // 0-based within the converted range
  const char* what;
};

struct Stats {
  int64_t lines = 0;    // lines looked at
  int64_t kept = 0;     // records written
  int64_t dropped = 0;  // records filtered out (language, missing/empty abstract)
  int64_t stopped = 0;  // 1 when an empty line ended the conversion
};

namespace detail {

constexpr int kMaxPosition = 1 << 24;

// Cursor over one line [p, e); reads past the end look like '\n', which no rule accepts.
struct Cur {
  const char* p;
  const char* e;
  char at() const { return p < e ? *p : '\n'; }
};

struct Bad {
  const char* what;
};

inline void ws(Cur& c) {
  while (c.p < c.e && (*c.p == ' ' || *c.p == '\t' || *c.p == '\r')) ++c.p;
}

inline void expect(Cur& c, char ch, const char* what) {
  if (c.at() != ch) throw Bad{what};
  ++c.p;
}

inline void skip_n(Cur& c, int n) {
  if (c.e - c.p < n) throw Bad{"truncated literal"};
  c.p += n;
}

// cursor on the opening quote -> raw (still escaped) contents; cursor after the closing quote
inline std::string_view raw_string(Cur& c) {
  expect(c, '"', "expected '\"'");
  const char* s = c.p;
  for (;;) {
    const char* q = static_cast<const char*>(memchr(c.p, '"', (size_t)(c.e - c.p)));
    if (!q) throw Bad{"unterminated string"};
    size_t backslashes = 0;
    for (const char* t = q; t > s && t[-1] == '\\'; --t) ++backslashes;
    c.p = q + 1;
    if ((backslashes & 1) == 0) return std::string_view(s, (size_t)(q - s));
  }
}

inline void open_composite(Cur& c, char ch) {
  ws(c);
  expect(c, ch, ch == '{' ? "expected '{'" : "expected '['");
  ws(c);
}

inline bool try_close(Cur& c, char ch) {
  if (c.at() != ch) return false;
  ++c.p;
  ws(c);
  return true;
}

inline void next_member(Cur& c) {
  if (c.at() == ',') ++c.p;
}

inline std::string_view name(Cur& c) {
  ws(c);
  std::string_view s = raw_string(c);
  ws(c);
  expect(c, ':', "expected ':'");
  return s;
}

inline void skip_value(Cur& c) {
  ws(c);
  const char ch = c.at();
  if ((ch >= '0' && ch <= '9') || ch == '-') {
    while (c.p < c.e) {
      const char x = *c.p;
      if ((x >= '0' && x <= '9') || x == '-' || x == '+' || x == 'e' || x == 'E' || x == '.') ++c.p;
      else break;
    }
  } else if (ch == 'f') {
    skip_n(c, 5);
  } else if (ch == 't' || ch == 'n') {
    skip_n(c, 4);
  } else if (ch == '"') {
    raw_string(c);
  } else if (ch == '{' || ch == '[') {
    int brackets = ch == '[', braces = ch == '{';
    ++c.p;
    while (brackets || braces) {
      if (c.p >= c.e) throw Bad{"unterminated array or object"};
      switch (*c.p) {
        case '"': raw_string(c); continue;
        case '[': ++brackets; break;
        case ']': --brackets; break;
        case '{': ++braces; break;
        case '}': --braces; break;
        default: break;
      }
      ++c.p;
    }
  } else {
    throw Bad{"not a JSON value"};
  }
  ws(c);
}

// "..." or null; returns false for null
inline bool nullable_string(Cur& c, std::string_view& out) {
  ws(c);
  bool have = false;
  if (c.at() == '"') {
    out = raw_string(c);
    have = true;
  } else if (c.at() == 'n') {
    skip_n(c, 4);
  } else {
    throw Bad{"expected string or null"};
  }
  ws(c);
  return have;
}

inline int position(Cur& c) {
  ws(c);
  if (c.at() == '-') throw Bad{"negative word position"};
  if (c.at() < '0' || c.at() > '9') throw Bad{"expected word position"};
  int64_t v = 0;
  while (c.p < c.e && *c.p >= '0' && *c.p <= '9') {
    v = v * 10 + (*c.p++ - '0');
    if (v > kMaxPosition) throw Bad{"word position too large"};
  }
  ws(c);
  return (int)v;
}

// Per-thread scratch reused across records.
struct Scratch {
  std::vector<std::string_view> words;  // slot -> word; data()==nullptr marks a gap
  std::string abstract;
};

// abstract_inverted_index value; returns false for null (abstract absent)
inline bool inverted_index(Cur& c, Scratch& s) {
  ws(c);
  if (c.at() == 'n') {
    skip_n(c, 4);
    ws(c);
    return false;
  }
  s.words.clear();
  for (open_composite(c, '{'); !try_close(c, '}'); next_member(c)) {
    const std::string_view word = name(c);
    for (open_composite(c, '['); !try_close(c, ']'); next_member(c)) {
      const int idx = position(c);
      if ((size_t)idx >= s.words.size()) s.words.resize((size_t)idx + 1);
      // an empty word must still count as placed: point it at the (non-null) key bytes
      s.words[(size_t)idx] = std::string_view(word.data(), word.size());
    }
  }
  s.abstract.clear();
  const size_t n = s.words.size();
  for (size_t i = 0; i < n; ++i) {
    if (s.words[i].data() == nullptr) continue;
    s.abstract.append(s.words[i].data(), s.words[i].size());
    if (i + 1 != n) s.abstract.push_back(' ');
  }
  return true;
}

// One record [b, e) (no newline) -> appended to out when kept.  Returns true when written.
inline bool convert_record(const char* b, const char* e, Scratch& s, std::string& out) {
  Cur c{b, e};
  std::string_view id, title;
  bool have_id = false, have_title = false, have_abstract = false;
  for (open_composite(c, '{'); !try_close(c, '}'); next_member(c)) {
    const std::string_view key = name(c);
    if (key == "id") {
      ws(c);
      id = raw_string(c);
      have_id = true;
      ws(c);
    } else if (key == "title") {
      have_title = nullable_string(c, title);
    } else if (key == "language") {
      std::string_view lang;
      if (!nullable_string(c, lang) || lang != "en") return false;
    } else if (key == "abstract_inverted_index") {
      have_abstract = inverted_index(c, s);
      if (!have_abstract || s.abstract.empty()) return false;
    } else {
      skip_value(c);
    }
  }
  if (!have_abstract) return false;
  out.append("{\"id\":\"");
  if (have_id) out.append(id.data(), id.size());
  else out.append("(null)");
  out.append("\",\"document\":\"");
  if (have_title) {
    out.append(title.data(), title.size());
    out.push_back(' ');
  }
  out.append(s.abstract);
  out.append("\"}\n");
  return true;
}

}  // namespace detail

// Converts the lines of [b, e) in order, appending to `out`.  Every line must be complete except
// that the last one may lack its '\n'.  Stops (stats.stopped = 1) at the first empty line.
inline void convert_lines(const char* b, const char* e, std::string& out, Stats& st) {
  detail::Scratch scratch;
  while (b < e) {
    const char* nl = static_cast<const char*>(memchr(b, '\n', (size_t)(e - b)));
    const char* le = nl ? nl : e;
    if (le == b) {  // empty line: the reference stops reading here
      st.stopped = 1;
      return;
    }
    try {
      if (detail::convert_record(b, le, scratch, out)) ++st.kept;
      else ++st.dropped;
    } catch (const detail::Bad& bad) {
      throw ParseError{st.lines, bad.what};
    }
    ++st.lines;
    b = nl ? nl + 1 : e;
  }
}

// Same, on `threads` threads: the range is cut at line boundaries into one piece per thread, the
// pieces are converted independently; `parts` receives their outputs in input order, cut off after
// the piece that met the first empty line (so concatenating `parts` equals convert_lines()).
inline void convert_lines_mt(const char* b, const char* e, int threads, std::vector<std::string>& parts,
                             Stats& st) {
  const size_t len = (size_t)(e - b);
  if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
  if (threads < 1) threads = 1;
  if ((size_t)threads > len / (1 << 16) + 1) threads = (int)(len / (1 << 16) + 1);  // >= 64 KB each
  parts.clear();
  if (threads == 1) {
    parts.emplace_back();
    convert_lines(b, e, parts.back(), st);
    return;
  }
  std::vector<const char*> cut((size_t)threads + 1);
  cut[0] = b;
  cut[(size_t)threads] = e;
  for (int t = 1; t < threads; ++t) {
    const char* guess = b + len * (size_t)t / (size_t)threads;
    if (guess < cut[(size_t)t - 1]) guess = cut[(size_t)t - 1];
    const char* nl = static_cast<const char*>(memchr(guess, '\n', (size_t)(e - guess)));
    cut[(size_t)t] = nl ? nl + 1 : e;
  }
  struct Piece {
    std::string out;
    Stats st;
    bool failed = false;
    ParseError err{0, nullptr};
  };
  std::vector<Piece> pieces((size_t)threads);
  std::vector<std::thread> pool;
  pool.reserve((size_t)threads);
  for (int t = 0; t < threads; ++t) {
    pool.emplace_back([&, t] {
      Piece& pc = pieces[(size_t)t];
      try {
        pc.out.reserve((size_t)(cut[(size_t)t + 1] - cut[(size_t)t]) / 2);
        convert_lines(cut[(size_t)t], cut[(size_t)t + 1], pc.out, pc.st);
      } catch (const ParseError& err) {
        pc.failed = true;
        pc.err = err;
      }
    });
  }
  for (auto& th : pool) th.join();
  for (auto& pc : pieces) {
    if (pc.failed) throw ParseError{st.lines + pc.err.line, pc.err.what};
    parts.emplace_back(std::move(pc.out));
    st.lines += pc.st.lines;
    st.kept += pc.st.kept;
    st.dropped += pc.st.dropped;
    if (pc.st.stopped) {
      st.stopped = 1;
      return;
    }
  }
}

// Copies `parts` back to back into dst (sum of their sizes), one thread per part.
inline void gather(const std::vector<std::string>& parts, char* dst) {
  std::vector<std::thread> pool;
  size_t off = 0;
  for (size_t i = 0; i < parts.size(); ++i) {
    const std::string& s = parts[i];
    char* to = dst + off;
    off += s.size();
    if (i + 1 == parts.size()) memcpy(to, s.data(), s.size());
    else pool.emplace_back([to, &s] { memcpy(to, s.data(), s.size()); });
  }
  for (auto& th : pool) th.join();
}

}  // namespace oa
}  // namespace absb
