// oa_jsonl.cu — C ABI of the OpenAlex JSON-lines front end (host-only code; see oa_jsonl.hpp).
// Replaces the `./oa_jsonl` pipeline stage, /root/reference/Makefile:64 (oa_jsonl.c:351-414).
#include <cstdlib>

#include "common.cuh"
#include "oa_jsonl.hpp"

using namespace absb;

extern "C" {

int absb_oa_jsonl_convert(const char* in, size_t in_len, int final_chunk, int threads, char** out,
                          size_t* out_len, size_t* consumed, int64_t* stats) {
  ABSB_API_BEGIN
  ABSB_CHECK(out && out_len && consumed, ABSB_ERR_INVALID, "null output argument");
  ABSB_CHECK(in || in_len == 0, ABSB_ERR_INVALID, "null input");
  *out = nullptr;
  *out_len = 0;
  *consumed = 0;
  // only complete lines are converted unless this is the last chunk of the stream
  size_t usable = in_len;
  if (!final_chunk) {
    while (usable > 0 && in[usable - 1] != '\n') --usable;
  }
  std::vector<std::string> parts;
  oa::Stats st;
  try {
    oa::convert_lines_mt(in, in + usable, threads, parts, st);
  } catch (const oa::ParseError& e) {
    fail(ABSB_ERR_INVALID, "oa_jsonl: malformed record on line %lld of this chunk: %s",
         (long long)(e.line + 1), e.what);
  }
  size_t total = 0;
  for (const auto& p : parts) total += p.size();
  char* buf = static_cast<char*>(malloc(total + 1));
  if (!buf) throw std::bad_alloc();
  oa::gather(parts, buf);
  buf[total] = '\0';
  *out = buf;
  *out_len = total;
  *consumed = st.stopped ? in_len : usable;
  if (stats) {
    stats[0] = st.lines;
    stats[1] = st.kept;
    stats[2] = st.dropped;
    stats[3] = st.stopped;
  }
  ABSB_API_END
}

int absb_oa_jsonl_free(char* out) {
  free(out);
  return ABSB_OK;
}

}  // extern "C"
