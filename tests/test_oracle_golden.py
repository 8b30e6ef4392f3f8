"""Oracle restatement vs committed golden vectors (tests/golden/ref_packets.json, produced by the unmodified
reference build via tests/golden/make_golden.py).  Needs neither /root/reference nor oracle/_ref: CPU only."""
import hashlib, json, os, pytest
from cases import CASES, make_frames
from oracle import ffv1_oracle as O, synth

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "ref_packets.json")))

def md5(b):
    return hashlib.md5(bytes(b)).hexdigest()

def check(g, frames):
    if md5(b"".join(f.tobytes() for f in frames)) != g["input_md5"]:
        pytest.skip("synthetic input differs from the fixture's (numpy version?)")
    o = O.Encoder(g["w"], g["h"], g["pix_fmt"], **g["opts"])
    assert md5(o.extradata) == g["extradata_md5"] and len(o.extradata) == g["extradata_size"]
    for i, f in enumerate(frames):
        pkt, key = o.encode(f)
        assert [len(pkt), md5(pkt), int(key)] == g["packets"][i], "packet %d" % i

@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_matrix(case):
    check(GOLD[case[0]], make_frames(case))

def test_s2_noisy1080_c2_first_gop_and_next_key():
    """BASELINE.json configs[1] at full size: 1080p yuv420p, GOP 16, coder=1, context=0, 24 slices; SURVEY App. B
    known answers (packet 0: 1 256 835 B 4d9b382c..., packet 16: 1 256 273 B b33944ce...)"""
    g = GOLD["s2_noisy1080_c2"]
    gen = synth.Noisy(1920, 1080, "yuv420p", 1234)
    frames = [gen.next() for _ in range(18)]
    o = O.Encoder(1920, 1080, "yuv420p", **g["opts"])
    assert md5(o.extradata) == "1f9c7f0a5eec5730d544c57ce84ad94f"
    for i, f in enumerate(frames):
        pkt, key = o.encode(f)
        assert [len(pkt), md5(pkt), int(key)] == g["packets"][i], "packet %d" % i
    assert g["packets"][0][:2] == [1256835, "4d9b382c959b10c2041e42bf442b8d0d"]
    assert g["packets"][16][:2] == [1256273, "b33944ce6f063654f55883d876f8360a"]
