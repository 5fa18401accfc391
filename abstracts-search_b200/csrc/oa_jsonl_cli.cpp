// oa_jsonl_cli.cpp — drop-in for the reference's `./oa_jsonl` executable (/root/reference/
// Makefile:64,68-69): OpenAlex works JSON-lines on stdin, {"id","document"} JSON-lines on stdout.
// Reads stdin in large blocks, converts the complete lines of each block on all host threads
// (oa_jsonl.hpp) and carries the unfinished tail over to the next block.
//   usage: oa_jsonl [threads]      (default: OA_JSONL_THREADS or all hardware threads)
#include <unistd.h>

#include <cstdio>
#include <cstdlib>

#include "oa_jsonl.hpp"

static bool write_all(const std::string& s) {
  size_t off = 0;
  while (off < s.size()) {
    const ssize_t w = write(STDOUT_FILENO, s.data() + off, s.size() - off);
    if (w <= 0) return false;
    off += (size_t)w;
  }
  return true;
}

int main(int argc, char** argv) {
  int threads = 0;
  if (const char* env = getenv("OA_JSONL_THREADS")) threads = atoi(env);
  if (argc > 1) threads = atoi(argv[1]);
  const size_t block = (size_t)32 << 20;
  std::string buf;
  std::vector<std::string> parts;
  buf.reserve(2 * block);
  absb::oa::Stats st;
  bool eof = false;
  while (!eof && !st.stopped) {
    const size_t have = buf.size();
    buf.resize(have + block);
    size_t got = 0;
    while (got < block) {
      const ssize_t r = read(STDIN_FILENO, &buf[have + got], block - got);
      if (r < 0) {
        perror("oa_jsonl: read");
        return 1;
      }
      if (r == 0) {
        eof = true;
        break;
      }
      got += (size_t)r;
    }
    buf.resize(have + got);
    size_t usable = buf.size();
    if (!eof) {
      while (usable > 0 && buf[usable - 1] != '\n') --usable;
    }
    try {
      absb::oa::convert_lines_mt(buf.data(), buf.data() + usable, threads, parts, st);
    } catch (const absb::oa::ParseError& e) {
      fprintf(stderr, "oa_jsonl: malformed record on line %lld: %s\n", (long long)(e.line + 1), e.what);
      return 2;
    }
    for (const auto& part : parts) {
      if (!write_all(part)) {
        perror("oa_jsonl: write");
        return 1;
      }
    }
    buf.erase(0, usable);
  }
  return 0;
}
