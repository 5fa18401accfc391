"""Generates the oa_jsonl golden vectors from the REFERENCE PROGRAM ITSELF: oracle/_ref/oa_jsonl,
compiled by oracle/Makefile from /root/reference/oa_jsonl.c (run in the build container; the
outputs are committed because /root/reference does not exist on the GPU box).

    python tests/golden/make_golden_oa.py
writes  tests/golden/oa_jsonl_cases.jsonl  -> oa_jsonl_cases.out   (hand-written edge cases, last
                                                                      line without '\\n')
        tests/golden/oa_jsonl_synth.jsonl  -> oa_jsonl_synth.out   (150 synthetic OpenAlex records)
        tests/golden/oa_jsonl_stop.jsonl   -> oa_jsonl_stop.out    (an empty line ends the run)
"""
import importlib
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import oa_jsonl as O  # noqa: E402

A = '"abstract_inverted_index":'
CASES = [
    # plain kept record; title + abstract
    '{"id":"W1","title":"T one","language":"en",' + A + '{"hello":[0],"world":[1]}}',
    # no title key / null title / empty title
    '{"id":"W2","language":"en",' + A + '{"solo":[0]}}',
    '{"id":"W3","title":null,"language":"en",' + A + '{"a":[0],"b":[1]}}',
    '{"id":"W4","title":"","language":"en",' + A + '{"a":[0],"b":[1]}}',
    # language: absent keeps, null / other / wrong case drop; language after the abstract still drops
    '{"id":"W5","title":"no language key",' + A + '{"kept":[0]}}',
    '{"id":"W6","title":"x","language":null,' + A + '{"a":[0]}}',
    '{"id":"W7","title":"x","language":"fr",' + A + '{"a":[0]}}',
    '{"id":"W8","title":"x","language":"EN",' + A + '{"a":[0]}}',
    '{"id":"W9","title":"x",' + A + '{"a":[0]},"language":"de"}',
    '{"id":"W10","title":"x",' + A + '{"a":[0]},"language":"en"}',
    # abstract: null, empty object, missing, only an empty word, empty word + real word
    '{"id":"W11","title":"x","language":"en",' + A + 'null}',
    '{"id":"W12","title":"x","language":"en",' + A + '{}}',
    '{"id":"W13","title":"x","language":"en"}',
    '{"id":"W14","title":"x","language":"en",' + A + '{"":[0]}}',
    '{"id":"W15","title":"x","language":"en",' + A + '{"":[0],"a":[1]}}',
    # gaps are skipped without a doubled space; the last slot never gets a trailing space
    '{"id":"W16","title":"gaps","language":"en",' + A + '{"a":[0],"d":[3],"h":[7]}}',
    '{"id":"W17","title":"late start","language":"en",' + A + '{"z":[5]}}',
    # a position claimed twice keeps the word parsed last; a word at many positions
    '{"id":"W18","title":"dup","language":"en",' + A + '{"a":[0,1],"b":[1]}}',
    '{"id":"W19","title":"dup","language":"en",' + A + '{"b":[1],"a":[0,1]}}',
    '{"id":"W20","title":"rep","language":"en",' + A + '{"the":[0,2,4],"cat":[1],"dog":[3],"end":[5]}}',
    # out-of-order words and positions beyond the reference's initial 100 slots (realloc path)
    '{"id":"W21","title":"big","language":"en",' + A + '{"far":[250],"near":[0],"mid":[120,121]}}',
    # missing id prints "(null)"; duplicate keys: the last one wins
    '{"title":"no id","language":"en",' + A + '{"a":[0]}}',
    '{"id":"first","id":"second","title":"t1","title":"t2","language":"en",' + A + '{"a":[0]}}',
    '{"id":"W24","language":"en",' + A + '{"old":[0]},' + A + '{"new":[0],"er":[1]}}',
    # escapes pass through untouched; escaped quote / backslash runs at string ends
    '{"id":"W25","title":"say \\"hi\\" \\\\","language":"en",' + A + '{"q\\"uote":[0],"back\\\\\\\\":[1],"\\u00e9t\\u00e9":[2],"\\ud83d\\ude00":[3]}}',
    '{"id":"W26","title":"\\\\\\"","language":"en",' + A + '{"tab\\there":[0],"nl\\nhere":[1]}}',
    # keys compare on raw bytes: an escaped spelling of "title" is just another key
    '{"id":"W27","t\\u0069tle":"not a title","language":"en",' + A + '{"a":[0]}}',
    # whitespace (space, tab, CR) wherever the grammar allows it
    ' \t{ "id" : "W28" ,\t"title"\t:\t"spaced" , "language" : "en" , ' + A + ' { "a" : [ 0 , 2 ] , "b" : [ 1 ] } } \r',
    '{"id":"W29","title":"crlf","language":"en",' + A + '{"a":[0]}}\r',
    # values that are skipped: nested composites with brackets inside strings, numbers, literals
    '{"x":{"a":"}]","b":[1,{"c":"\\"]"}],"d":{}},"id":"W30","n":-1.5e+3,"t":true,"f":false,"z":null,"l":[],"s":"","title":"skips","language":"en","arr":[[["deep"]]],' + A + '{"ok":[0]}}',
    '{"authorships":[{"author":{"display_name":"A \\"B\\" {C} [D]"},"raw":["x\\\\"]}],"id":"W31","title":"more skips","language":"en",' + A + '{"ok":[0]},"counts_by_year":[{"year":2020,"cited_by_count":1}]}',
    # anything after the closing brace is ignored; an empty object prints nothing
    '{"id":"W32","title":"trailing","language":"en",' + A + '{"a":[0]}} trailing garbage {',
    '{}',
    # UTF-8 passes through
    '{"id":"W34","title":"Überraschung — naïve café","language":"en",' + A + '{"日本語":[0],"ελληνικά":[1]}}',
    # very long word list
    '{"id":"W35","title":"long","language":"en",' + A + '{' + ",".join(f'"w{i}":[{i}]' for i in range(400)) + '}}',
    # commas are optional to the reference's grammar (parse_*_next only skips one if present)
    '{"id":"W37" "title":"missing commas" "language":"en" ' + A + '{"a":[0] "b":[1 2]}}',
    # last line has no newline
    '{"id":"W36","title":"no newline at end","language":"en",' + A + '{"fin":[0]}}',
]


def main() -> None:
    assert O.reference_available(), "build oracle/_ref/oa_jsonl first: make -C oracle"
    P = importlib.import_module("abstracts-search_b200")
    cases = "\n".join(CASES).encode("utf-8")
    synth = P.oa_jsonl.synth_records(20240501, 150, mean_words=60, filler=4)
    stop = synth[: synth.index(b"\n", len(synth) // 3) + 1] + b"\n" + synth[len(synth) // 2:]
    for name, data in (("cases", cases), ("synth", synth), ("stop", stop)):
        out = O.convert_reference(data)
        open(os.path.join(HERE, f"oa_jsonl_{name}.jsonl"), "wb").write(data)
        open(os.path.join(HERE, f"oa_jsonl_{name}.out"), "wb").write(out)
        print(name, len(data), "->", len(out), "bytes,", out.count(b"\n"), "records")


if __name__ == "__main__":
    main()
