"""OpenAlex JSON-lines front end: product (libabsb200.so C ABI + the oa_jsonl
executable) vs. the golden vectors made by the reference program, vs. the Python restatement in
oracle/, and — where oracle/_ref/oa_jsonl is present (built from /root/reference/oa_jsonl.c) —
vs. the reference program itself on fresh seeded input.  Bar: byte-exact."""
import io
import json
import os
import subprocess

import pytest

from conftest import GOLDEN, ROOT, load_pkg
from oracle import oa_jsonl as O

P = load_pkg()
OA = P.oa_jsonl
CLI = os.path.join(ROOT, "abstracts-search_b200", "oa_jsonl")
NAMES = ("cases", "synth", "stop")


def _golden(name):
    return (open(os.path.join(GOLDEN, f"oa_jsonl_{name}.jsonl"), "rb").read(),
            open(os.path.join(GOLDEN, f"oa_jsonl_{name}.out"), "rb").read())


@pytest.mark.parametrize("name", NAMES)
def test_oracle_restatement_matches_reference_golden(name):
    data, want = _golden(name)
    assert O.convert(data) == want


@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("threads", [1, 3, 8])
def test_product_matches_reference_golden(name, threads):
    data, want = _golden(name)
    st = OA.ConvertStats()
    assert OA.convert(data, threads=threads, stats=st) == want
    assert st["kept"] == want.count(b"\n")
    assert st["stopped"] == (1 if name == "stop" else 0)


@pytest.mark.parametrize("name", NAMES)
def test_cli_matches_reference_golden(name):
    data, want = _golden(name)
    r = subprocess.run([CLI, "2"], input=data, capture_output=True, timeout=60)
    assert r.returncode == 0 and r.stdout == want


def test_output_is_json_the_next_stage_can_parse():
    data, want = _golden("cases")
    docs = dict(OA.iter_documents(io.BytesIO(data), threads=2, block_bytes=777))
    assert docs["W25"] == 'say "hi" \\ q"uote back\\\\ été 😀'
    assert docs["W16"] == "gaps a d h" and docs["W4"] == " a b" and docs["(null)"] == "no id a"
    assert len(docs) == want.count(b"\n")
    for line in want.splitlines():
        json.loads(line)


@pytest.mark.parametrize("block", [1 << 10, 5000, 1 << 16])
def test_streaming_in_blocks_equals_one_shot(block):
    for name in NAMES:
        data, want = _golden(name)
        dst = io.BytesIO()
        st = OA.convert_stream(io.BytesIO(data), dst, threads=2, block_bytes=block)
        assert dst.getvalue() == want
        assert st["kept"] == want.count(b"\n")


def test_partial_chunk_leaves_the_unfinished_line():
    data, _ = _golden("synth")
    cut = data.index(b"\n", len(data) // 2) + 1
    head, used = OA.convert_partial(data[: cut + 100], final=False)
    assert used == cut
    rest = OA.convert(data[cut:], final=True)
    assert head + rest == OA.convert(data)


def test_empty_and_degenerate_inputs():
    assert OA.convert(b"") == b""
    assert OA.convert(b"\n") == b""
    assert OA.convert(b"{}\n") == b""
    assert OA.convert(b"{}") == b""
    assert OA.convert(b'{"id":"a"', final=False) == b""  # no complete line yet


@pytest.mark.parametrize("bad", [
    b'{"id":"W1","title":"unterminated\n',
    b'["not an object"]\n',
    b'{"id":"W1","abstract_inverted_index":{"a":[-1]}}\n',
    b'{"id":"W1","abstract_inverted_index":{"a":[99999999999]}}\n',
    b'{"id":"W1","x":[1,2\n',
    b' \r\n',
])
def test_malformed_records_raise_instead_of_asserting(bad):
    good = b'{"id":"ok","language":"en","abstract_inverted_index":{"a":[0]}}\n'
    with pytest.raises(P.AbsbError) as e:
        OA.convert(good + bad, threads=1)
    assert "line 2" in str(e.value)
    with pytest.raises(O.Malformed):
        O.convert(good + bad)
    r = subprocess.run([CLI, "1"], input=good + bad, capture_output=True, timeout=60)
    assert r.returncode == 2 and b"line 2" in r.stderr


@pytest.mark.skipif(not O.reference_available(), reason="oracle/_ref/oa_jsonl not built (needs /root/reference)")
@pytest.mark.parametrize("seed,n,words,filler", [(1, 400, 180, 12), (2, 800, 20, 2), (3, 100, 900, 30), (4, 2000, 5, 1)])
def test_product_and_oracle_match_the_reference_program_on_fresh_input(seed, n, words, filler):
    data = OA.synth_records(seed, n, mean_words=words, filler=filler)
    want = O.convert_reference(data)
    assert want.count(b"\n") > n // 4
    assert OA.convert(data, threads=1) == want
    assert OA.convert(data, threads=8) == want
    assert O.convert(data) == want
    assert OA.convert(data[:-1], threads=4) == want  # final line without '\n'


def test_threads_preserve_record_order_on_a_large_block():
    data = OA.synth_records(9, 3000, mean_words=40, filler=3)
    one = OA.convert(data, threads=1)
    assert OA.convert(data, threads=16) == one
    ids = [json.loads(x)["id"] for x in one.splitlines()]
    src = [json.loads(x)["id"] for x in data.splitlines()]
    assert ids == [i for i in src if i in set(ids)]


@pytest.mark.skipif(not O.reference_available(), reason="oracle/_ref/oa_jsonl not built (needs /root/reference)")
def test_cli_carries_partial_lines_across_its_32mb_blocks():
    """The executable reads stdin in 32 MB blocks and carries the unfinished tail over: an input of
    several blocks must come out byte-identical to the reference program's output."""
    data = OA.synth_records(5, 11000, mean_words=120, filler=10)
    assert len(data) > 36 << 20
    want = O.convert_reference(data)
    r = subprocess.run([CLI], input=data, capture_output=True, timeout=120)
    assert r.returncode == 0 and r.stdout == want
    r1 = subprocess.run([CLI, "1"], input=data[:-1], capture_output=True, timeout=120)  # no final newline, one thread
    assert r1.returncode == 0 and r1.stdout == want


# ---------------------------------------------------------------------------------------------
# property-based: arbitrary well-formed records (json.dumps of random nested values) through the
# reference program, the restatement and the product
# ---------------------------------------------------------------------------------------------
try:
    from hypothesis import HealthCheck, given, settings
    from hypothesis import strategies as hst

    _leaf = hst.one_of(hst.none(), hst.booleans(), hst.integers(-10**9, 10**9),
                       hst.floats(allow_nan=False, allow_infinity=False, width=64),
                       hst.text(max_size=12))
    _json = hst.recursive(_leaf, lambda ch: hst.one_of(hst.lists(ch, max_size=4),
                                                       hst.dictionaries(hst.text(max_size=6), ch, max_size=4)), max_leaves=12)
    _words = hst.dictionaries(hst.text(max_size=8), hst.lists(hst.integers(0, 150), max_size=4), max_size=12)

    @hst.composite
    def _record(draw):
        members = []
        for i in range(draw(hst.integers(0, 4))):
            members.append((draw(hst.text(min_size=1, max_size=6)) + str(i), draw(_json)))
        if draw(hst.booleans()):
            members.append(("id", draw(hst.text(max_size=20))))
        if draw(hst.booleans()):
            members.append(("title", draw(hst.one_of(hst.none(), hst.text(max_size=30)))))
        if draw(hst.booleans()):
            members.append(("language", draw(hst.one_of(hst.none(), hst.sampled_from(["en", "en", "fr", "EN", ""])))))
        if draw(hst.booleans()):
            members.append(("abstract_inverted_index", draw(hst.one_of(hst.none(), _words))))
        members = draw(hst.permutations(members))
        rec = {}
        for key, val in members:
            if key not in rec:
                rec[key] = val
        seps = draw(hst.sampled_from([(",", ":"), (", ", ": "), (" ,\t", " :\t")]))
        return json.dumps(rec, separators=seps, ensure_ascii=draw(hst.booleans()))

    @pytest.mark.skipif(not O.reference_available(), reason="oracle/_ref/oa_jsonl not built (needs /root/reference)")
    @settings(max_examples=150, deadline=None, suppress_health_check=[HealthCheck.too_slow])
    @given(hst.lists(_record(), min_size=1, max_size=12), hst.booleans())
    def test_random_wellformed_records_match_the_reference_program(records, final_newline):
        data = ("\\n".join(records) + ("\\n" if final_newline else "")).encode("utf-8")
        want = O.convert_reference(data)
        assert OA.convert(data, threads=1) == want
        assert OA.convert(data, threads=3) == want
        assert O.convert(data) == want
except ImportError:  # hypothesis is optional
    pass
