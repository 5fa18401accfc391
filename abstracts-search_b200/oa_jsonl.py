"""OpenAlex `works` JSON-lines -> {"id","document"} JSON-lines: the stage in front of bulk encode.

Host-side mirror of the reference's `./oa_jsonl` pipeline stage (/root/reference/Makefile:64,
program /root/reference/oa_jsonl.c:351-414) over `absb_oa_jsonl_convert` (csrc/oa_jsonl.hpp):
same bytes out on well-formed input, but block-oriented and multi-threaded by line range instead
of fgetc-at-a-time.  `abstracts-search_b200/oa_jsonl` (built by csrc/Makefile) is the stdin->stdout
executable that drops into the Makefile pipeline unchanged.  SURVEY §8f row 4.
"""
from __future__ import annotations

import ctypes
import json
from typing import BinaryIO, Iterator

import numpy as np

from ._lib import check, lib


class ConvertStats(dict):
    """lines / kept / dropped / stopped counters of one convert() call."""


def convert(data: bytes, threads: int = 0, final: bool = True, stats: ConvertStats | None = None) -> bytes:
    """Convert a block of OpenAlex JSONL.  With final=False only the complete lines are converted
    (use `convert_partial` to learn how many bytes were consumed)."""
    out, _ = convert_partial(data, threads=threads, final=final, stats=stats)
    return out


def convert_partial(data: bytes, threads: int = 0, final: bool = False,
                    stats: ConvertStats | None = None) -> tuple[bytes, int]:
    """-> (converted bytes, number of input bytes consumed).  A malformed record raises
    AbsbError(ERR_INVALID) naming the line (the reference aborts on an assert instead)."""
    buf = ctypes.c_char_p()
    out_len, consumed = ctypes.c_size_t(0), ctypes.c_size_t(0)
    st = (ctypes.c_int64 * 4)()
    L = lib()
    check(L.absb_oa_jsonl_convert(data, len(data), int(final), int(threads), ctypes.byref(buf),
                                  ctypes.byref(out_len), ctypes.byref(consumed), st))
    try:
        out = ctypes.string_at(buf, out_len.value)
    finally:
        L.absb_oa_jsonl_free(buf)
    if stats is not None:
        stats.update(lines=st[0], kept=st[1], dropped=st[2], stopped=st[3])
    return out, consumed.value


def convert_stream(src: BinaryIO, dst: BinaryIO, threads: int = 0, block_bytes: int = 32 << 20) -> ConvertStats:
    """stdin->stdout behaviour of the executable, for file objects."""
    total = ConvertStats(lines=0, kept=0, dropped=0, stopped=0)
    tail = b""
    while True:
        block = src.read(block_bytes)
        eof = not block
        data = tail + block
        st = ConvertStats()
        out, used = convert_partial(data, threads=threads, final=eof, stats=st)
        dst.write(out)
        for key in ("lines", "kept", "dropped"):
            total[key] += st[key]
        if st["stopped"]:
            total["stopped"] = 1
            break
        tail = data[used:]
        if eof:
            break
    return total


def iter_documents(src: BinaryIO, threads: int = 0, block_bytes: int = 32 << 20) -> Iterator[tuple[str, str]]:
    """(id, document) pairs as `sidecar-search build` reads them from the pipe (Makefile:64-65):
    the converter's output lines parsed as JSON, escapes resolved there."""
    tail = b""
    while True:
        block = src.read(block_bytes)
        eof = not block
        data = tail + block
        st = ConvertStats()
        out, used = convert_partial(data, threads=threads, final=eof, stats=st)
        for line in out.splitlines():
            rec = json.loads(line)
            yield rec["id"], rec["document"]
        if st["stopped"] or eof:
            return
        tail = data[used:]


# ------------------------------------------------------------------------------------------------
# synthetic OpenAlex `works` records (the reference ships no data and there is no network)
# ------------------------------------------------------------------------------------------------
_VOCAB = ("the of and in to a is for with that by on as are from this we be an which at results "
          "model data using method analysis study based between effect system two these can has "
          "were new different high protein cell quantum graph neural energy learning surface "
          "temperature patients clinical theory field structure dynamics optimal \\\"quoted\\\" "
          "caf\\u00e9 na\\u00efve \\ud83d\\ude00 back\\\\slash tab\\there \\u03b1-helix 10\\u00b0C "
          "p<0.05 [1] {x} a,b").split(" ")
_LANGS = ("en", "en", "en", "en", "en", "en", "en", "fr", "de", "zh", "es")


def synth_records(seed: int, n: int, mean_words: int = 180, filler: int = 12) -> bytes:
    """n OpenAlex-shaped records as JSON lines: the four keys the converter looks at among the
    nested arrays/objects it has to skip (authorships, concepts, referenced_works ...), with the
    edge cases the reference's logic distinguishes: null / missing / non-English language, null
    and empty abstract_inverted_index, null and empty titles, position gaps, repeated words,
    escaped quotes and backslashes, \\u escapes incl. surrogate pairs, loose whitespace."""
    rng = np.random.default_rng(seed)
    lines = []
    for r in range(n):
        sp = " " if rng.random() < 0.3 else ""
        wid = int(rng.integers(1, 1 << 40))
        members = []
        members.append(f'"id":{sp}"https://openalex.org/W{wid}"')
        members.append(f'"doi":{sp}' + ("null" if rng.random() < 0.2 else f'"https://doi.org/10.{wid % 9999}/x{r}"'))
        u = rng.random()
        nt = int(rng.integers(3, 16))
        title_words = [_VOCAB[int(i)] for i in rng.integers(0, len(_VOCAB), nt)]
        if u < 0.05:
            members.append(f'"title":{sp}null')
        elif u < 0.07:
            members.append(f'"title":{sp}""')
        else:
            members.append(f'"title":{sp}"' + " ".join(title_words) + '"')
        members.append(f'"display_name":{sp}"' + " ".join(title_words) + '"')
        members.append(f'"publication_year":{sp}{int(rng.integers(1900, 2025))}')
        u = rng.random()
        lang_member = None
        if u < 0.05:
            lang_member = f'"language":{sp}null'
        elif u < 0.93:
            lang_member = f'"language":{sp}"{_LANGS[int(rng.integers(0, len(_LANGS)))]}"'
        if lang_member:
            members.append(lang_member)
        members.append(f'"type":{sp}"article"')
        members.append(f'"is_retracted":{sp}' + ("true" if rng.random() < 0.01 else "false"))
        members.append(f'"cited_by_count":{sp}{int(rng.integers(0, 5000))}')
        members.append(f'"fwci":{sp}{rng.random() * 10:.3f}')
        members.append(f'"apc_paid":{sp}' + ("null" if rng.random() < 0.7 else '{"value":1.5e3,"currency":"USD"}'))
        auth = []
        for a in range(int(rng.integers(1, filler + 1))):
            auth.append('{"author_position":"middle","author":{"id":"https://openalex.org/A%d","display_name":"A. %s [%d] \\"Q\\" }{"},'
                        '"institutions":[{"id":"https://openalex.org/I%d","country_code":"US","lineage":["https://openalex.org/I%d"]}],'
                        '"is_corresponding":false,"raw_affiliation_strings":["Dept. of X, Univ. \\\\ Y"]}'
                        % (int(rng.integers(1, 1 << 32)), _VOCAB[int(rng.integers(0, 40))], a,
                           int(rng.integers(1, 1 << 32)), int(rng.integers(1, 1 << 32))))
        members.append(f'"authorships":{sp}[' + ",".join(auth) + "]")
        members.append(f'"concepts":{sp}[' + ",".join(
            '{"id":"https://openalex.org/C%d","level":%d,"score":%.4f}' % (int(rng.integers(1, 1 << 30)), int(rng.integers(0, 5)), rng.random())
            for _ in range(int(rng.integers(0, filler + 1)))) + "]")
        members.append(f'"referenced_works":{sp}[' + ",".join(
            '"https://openalex.org/W%d"' % int(rng.integers(1, 1 << 40)) for _ in range(int(rng.integers(0, 4 * filler)))) + "]")
        u = rng.random()
        if u < 0.2:
            aii = f'"abstract_inverted_index":{sp}null'
        elif u < 0.22:
            aii = f'"abstract_inverted_index":{sp}{{}}'
        elif u < 0.24:
            aii = None
        else:
            nw = max(1, int(rng.poisson(mean_words)))
            picks = rng.integers(0, len(_VOCAB), nw)
            inv: dict[str, list[int]] = {}
            pos = 0
            for wi in picks:
                if rng.random() < 0.01:
                    pos += int(rng.integers(1, 4))  # a gap: positions nobody claims
                inv.setdefault(_VOCAB[int(wi)], []).append(pos)
                pos += 1
            items = list(inv.items())
            order = rng.permutation(len(items))
            aii = f'"abstract_inverted_index":{sp}{{' + f",{sp}".join(
                f'"{items[int(i)][0]}":{sp}[' + f",{sp}".join(str(p) for p in items[int(i)][1]) + "]" for i in order) + "}"
        if aii:
            members.append(aii)
        members.append(f'"counts_by_year":{sp}[{{"year":2023,"cited_by_count":3}},{{"year":2022,"cited_by_count":-1}}]')
        members.append(f'"updated_date":{sp}"2024-05-0{1 + r % 9}T00:00:00.000000"')
        if rng.random() < 0.1:  # key order is not fixed: move language (if any) to the very end
            if lang_member:
                members.remove(lang_member)
                members.append(lang_member)
        lines.append("{" + f",{sp}".join(members) + "}")
    return ("\n".join(lines) + "\n").encode("utf-8")
