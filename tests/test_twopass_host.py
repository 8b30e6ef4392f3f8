"""Two-pass coding, host half (no GPU): stats_in -> sorted state-transition table and initial states -> extradata
(ffv1enc.c:906-986, 621-667, 139-183, 591-606) must equal what the reference build published for the same statistics.
The statistics and the expected extradata come from tests/golden/twopass.json (generated from oracle/_ref by
tests/golden/make_twopass_golden.py); with the reference build present the fixture itself is re-derived as well."""
import base64, json, os, zlib, pytest
from twopass_cases import TWOPASS_CASES, make_frames

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "twopass.json")))

def stats_of(cid):
    return zlib.decompress(base64.b64decode(GOLD[cid]["stats_z"])).decode()

def gpu_opts(opts):
    o = dict(opts)
    o["g"] = o.pop("gop")
    return o

@pytest.mark.parametrize("case", TWOPASS_CASES, ids=[c[0] for c in TWOPASS_CASES])
def test_second_pass_extradata_matches_reference(case):
    import ffv1_b200
    cid, w, h, fmt, opts, n = case
    info, xd = ffv1_b200.resolve_encoder(w, h, fmt, flags=ffv1_b200.FLAG_PASS2, stats_in=stats_of(cid), **gpu_opts(opts))
    assert xd.hex() == GOLD[cid]["pass2_extradata"]
    info, xd1 = ffv1_b200.resolve_encoder(w, h, fmt, flags=ffv1_b200.FLAG_PASS1, **gpu_opts(opts))
    assert xd1.hex() == GOLD[cid]["pass1_extradata"] and info.version == 3      # the pass flags lift the version to 2 -> 3

def test_broken_statistics_are_refused():
    import ffv1_b200
    with pytest.raises(ffv1_b200.FFV1Error) as e:
        ffv1_b200.resolve_encoder(176, 144, "yuv420p", g=3, level=3, coder=1, slices=4, flags=ffv1_b200.FLAG_PASS2, stats_in="1 2 3 oops")
    assert e.value.code == -1094995529                                          # AVERROR_INVALIDDATA, ffv1enc.c:920-924
    with pytest.raises(ffv1_b200.FFV1Error):                                    # version 0/1 cannot carry initial states
        ffv1_b200.resolve_encoder(96, 80, "yuv420p", g=3, level=1, coder=1, stats_in=stats_of("tp_420_custom"))

def test_fixture_is_what_the_reference_build_produces(ref):
    import hashlib
    case = TWOPASS_CASES[0]
    cid, w, h, fmt, opts, n = case
    frames = make_frames(case)
    e1 = ref.Encoder(w, h, fmt, two_pass=1, **opts)
    for f in frames:
        e1.encode(f)
    st = e1.stats_out()
    assert hashlib.md5(st.encode()).hexdigest() == GOLD[cid]["stats_md5"]
    e2 = ref.Encoder(w, h, fmt, two_pass=2, stats_in=st, **opts)
    assert e2.extradata.hex() == GOLD[cid]["pass2_extradata"]
