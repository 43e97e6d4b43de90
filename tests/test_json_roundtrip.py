"""aeson-compatible JSON of the host mirror: Searcher (Searcher.hs:68-77), Replacer / Payload
(generic instances, Replacer.hs:56-83), Splitter (Splitter.hs:54-60).  Host images only (device = -2): runs on the CPU."""
import json

import pytest


def test_searcher_json_roundtrip():
    from alfred_margaret_b200 import searcher as S
    from alfred_margaret_b200.case_sensitivity import CaseSensitive, IgnoreCase
    s = S.build(CaseSensitive, ["tshirt", "shirts", "shorts"], device=-2)
    j = S.to_json(s)
    assert j == {"needles": [["tshirt", []], ["shirts", []], ["shorts", []]], "caseSensitivity": "CaseSensitive"}   # () encodes as []
    back = S.from_json(json.loads(json.dumps(j)), device=-2)
    assert back == s and S.num_needles(back) == 3 and S.case_sensitivity(back) == CaseSensitive
    sv = S.build_with_values(IgnoreCase, [("groß", 7), ("öffnung", 8)], device=-2)
    jv = S.to_json(sv)
    assert jv["caseSensitivity"] == "IgnoreCase" and jv["needles"] == [["groß", 7], ["öffnung", 8]]
    assert S.from_json(json.loads(json.dumps(jv, ensure_ascii=False)), device=-2) == sv
    with pytest.raises(ValueError):
        S.from_json({"needles": []}, device=-2)
    with pytest.raises(KeyError):
        S.from_json({"needles": [], "caseSensitivity": "Sometimes"}, device=-2)


def test_replacer_json_roundtrip():
    from alfred_margaret_b200 import replacer as R
    from alfred_margaret_b200.case_sensitivity import CaseSensitive, IgnoreCase
    r = R.build(IgnoreCase, [("Éclair", "lightning"), ("ẞèta", "sseta"), ("foo", "")], device=-2)
    j = R.to_json(r)
    needles = j["replacerSearcher"]["needles"]
    assert j["replacerSearcher"]["caseSensitivity"] == "IgnoreCase"
    # stored needles are lowered (Replacer.hs:105-107); lengths are those of the ORIGINAL needle (:111-113); priority -i
    assert [n for n, _ in needles] == ["éclair", "ßèta", "foo"]
    assert needles[1][1] == {"needlePriority": -1, "needleLengthBytes": len("ẞèta".encode()), "needleLengthCodePoints": 4, "needleReplacement": "sseta"}
    back = R.from_json(json.loads(json.dumps(j, ensure_ascii=False)), device=-2)
    assert R.to_json(back)["replacerSearcher"]["needles"][0][0] == "éclair"
    assert [p["needleReplacement"] for _, p in R.to_json(back)["replacerSearcher"]["needles"]] == ["lightning", "sseta", ""]
    rc = R.build(CaseSensitive, [("A", "B"), ("X", "Y")], device=-2)
    assert R.from_json(R.to_json(rc), device=-2) == rc
    bad = R.to_json(rc)
    bad["replacerSearcher"]["needles"][1][1]["needlePriority"] = 0
    with pytest.raises(ValueError):
        R.from_json(bad, device=-2)


def test_splitter_json_roundtrip():
    from alfred_margaret_b200 import splitter as Sp
    s = Sp.build(", ", device=-2)
    assert Sp.to_json(s) == ", " and Sp.from_json(json.loads(json.dumps(Sp.to_json(s))), device=-2) == s
    with pytest.raises(ValueError):
        Sp.from_json(["not", "a", "string"])
