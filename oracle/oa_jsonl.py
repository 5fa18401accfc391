"""CPU oracle (TEST INFRASTRUCTURE ONLY) for the OpenAlex JSON-lines front end, SURVEY §8f row 4.

Pure-Python restatement of /root/reference/oa_jsonl.c, function by function; only `tests/`,
`__graft_entry__.smoke()` and bench.py's CPU legs may import it.  Parity status: PINNED — the
reference program itself compiles here (oracle/Makefile -> oracle/_ref/oa_jsonl, straight from
/root/reference/oa_jsonl.c) and tests/test_oa_jsonl_cpu.py checks this restatement, the committed
golden vectors (tests/golden/oa_jsonl_*.{jsonl,out}, made by tests/golden/make_golden_oa.py from
that binary) and the product (libabsb200.so / the oa_jsonl CLI) against it byte for byte.
"""
from __future__ import annotations

import os
import subprocess

REF_BIN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "oa_jsonl")

_WS = b" \t\r"  # advance_space, oa_jsonl.c:43-47 ('\n' is not whitespace in JSONL)
_NUM0 = b"0123456789-"  # initial_number_char :35-37
_NUM = _NUM0 + b"+eE."  # number_char :39-41


class Malformed(ValueError):
    """The reference assert()s (aborts) on these inputs."""


def _space(s: bytes, i: int) -> int:
    while i < len(s) and s[i] in _WS:
        i += 1
    return i


def _string(s: bytes, i: int) -> tuple[bytes, int]:
    """advance_string :49-71: the closing quote is the first '"' preceded by an even number of
    backslashes.  Returns (raw contents, index after the closing quote)."""
    if i >= len(s) or s[i] != 0x22:
        raise Malformed("expected string")
    start = i + 1
    j = start
    while True:
        j = s.find(b'"', j)
        if j < 0:
            raise Malformed("unterminated string")
        cnt, t = 0, j - 1
        while t >= start and s[t] == 0x5C:
            cnt += 1
            t -= 1
        if cnt % 2 == 0:
            return s[start:j], j + 1
        j += 1


def _open(s: bytes, i: int, ch: int) -> int:  # advance_composite_open :73-82
    i = _space(s, i)
    if i >= len(s) or s[i] != ch:
        raise Malformed("expected opening bracket")
    return _space(s, i + 1)


def _try_close(s: bytes, i: int, ch: int):  # advance_composite_try_close :84-91
    if i < len(s) and s[i] == ch:
        return _space(s, i + 1)
    return None


def _next(s: bytes, i: int) -> int:  # advance_composite_next :93-98
    return i + 1 if i < len(s) and s[i] == 0x2C else i


def _skip_value(s: bytes, i: int) -> int:  # advance_value_skip :100-133
    i = _space(s, i)
    if i >= len(s):
        raise Malformed("missing value")
    c = s[i]
    if c in _NUM0:
        while i < len(s) and s[i] in _NUM:
            i += 1
    elif c == ord("f"):
        i += 5
    elif c in (ord("t"), ord("n")):
        i += 4
    elif c == 0x22:
        _, i = _string(s, i)
    elif c in (ord("{"), ord("[")):
        brackets, braces = int(c == ord("[")), int(c == ord("{"))
        i += 1
        while brackets or braces:
            if i >= len(s):
                raise Malformed("unterminated composite")
            c = s[i]
            if c == 0x22:
                _, i = _string(s, i)
                continue
            brackets += (c == ord("[")) - (c == ord("]"))
            braces += (c == ord("{")) - (c == ord("}"))
            i += 1
    else:
        raise Malformed("not a JSON value")
    if i > len(s):
        raise Malformed("truncated literal")
    return _space(s, i)


def _name(s: bytes, i: int) -> tuple[bytes, int]:  # parse_name :160-165
    i = _space(s, i)
    key, i = _string(s, i)
    i = _space(s, i)
    if i >= len(s) or s[i] != ord(":"):
        raise Malformed("expected ':'")
    return key, i + 1


def _nullable_string(s: bytes, i: int):  # parse_nullable_string :143-158
    i = _space(s, i)
    if i < len(s) and s[i] == 0x22:
        v, i = _string(s, i)
    elif i < len(s) and s[i] == ord("n"):
        v, i = None, i + 4
    else:
        raise Malformed("expected string or null")
    return v, _space(s, i)


def _abstract(s: bytes, i: int):
    """oajsonl_parse_abstract_inverted_index :284-325 + oajsonl_add_word :232-250 +
    oajsonl_build_abstract :260-282."""
    i = _space(s, i)
    if i < len(s) and s[i] == ord("n"):
        return None, _space(s, i + 4)
    words: list = []
    i = _open(s, i, ord("{"))
    while True:
        j = _try_close(s, i, ord("}"))
        if j is not None:
            i = j
            break
        word, i = _name(s, i)
        i = _open(s, i, ord("["))
        while True:
            j = _try_close(s, i, ord("]"))
            if j is not None:
                i = j
                break
            i = _space(s, i)
            k = i
            while k < len(s) and 0x30 <= s[k] <= 0x39:
                k += 1
            if k == i:
                raise Malformed("expected a non-negative word position")
            idx = int(s[i:k])
            if idx > 1 << 24:
                raise Malformed("word position too large (the reference would realloc gigabytes)")
            i = _space(s, k)
            if idx >= len(words):
                words.extend([None] * (idx + 1 - len(words)))
            words[idx] = word
            i = _next(s, i)
        i = _next(s, i)
    out = bytearray()
    for n, w in enumerate(words):
        if w is None:
            continue
        out += w
        if n != len(words) - 1:
            out += b" "
    return bytes(out), i


def convert_line(line: bytes):
    """One record (without its newline) -> output line, or None when the record is dropped
    (main loop, oa_jsonl.c:360-411)."""
    rid = title = abstract = None
    i = _open(line, 0, ord("{"))
    while True:
        if _try_close(line, i, ord("}")) is not None:
            break
        key, i = _name(line, i)
        if key == b"id":
            i = _space(line, i)
            rid, i = _string(line, i)
            i = _space(line, i)
        elif key == b"title":
            title, i = _nullable_string(line, i)
        elif key == b"language":
            lang, i = _nullable_string(line, i)
            if lang != b"en":
                return None
        elif key == b"abstract_inverted_index":
            abstract, i = _abstract(line, i)
            if not abstract:
                return None
        else:
            i = _skip_value(line, i)
        i = _next(line, i)
    if abstract is None:
        return None
    rid = b"(null)" if rid is None else rid  # glibc printf("%s", NULL)
    doc = abstract if title is None else title + b" " + abstract
    return b'{"id":"' + rid + b'","document":"' + doc + b'"}\n'


def convert(data: bytes) -> bytes:
    """Whole stream: lines end at '\\n' or end of input (read_line :333-349); the first empty line
    ends the conversion (:363-366)."""
    out = []
    for line in data.split(b"\n"):
        if not line:
            break
        r = convert_line(line)
        if r is not None:
            out.append(r)
    return b"".join(out)


def reference_available() -> bool:
    return os.access(REF_BIN, os.X_OK)


def convert_reference(data: bytes, timeout: float = 600.0) -> bytes:
    """The reference program itself (oracle/_ref/oa_jsonl, compiled from /root/reference)."""
    r = subprocess.run([REF_BIN], input=data, capture_output=True, timeout=timeout)
    if r.returncode != 0:
        raise Malformed(f"reference oa_jsonl exited {r.returncode}: {r.stderr[-200:]!r}")
    return r.stdout
