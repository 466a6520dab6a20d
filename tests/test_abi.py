"""The C-ABI library loads and exports every symbol include/prl.h declares; World ingest follows the
serde schema of the reference (src/common.rs:19-49); nothing here needs a GPU."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest

import physics_reinforcement_learning_environment_b200 as prl
from physics_reinforcement_learning_environment_b200 import _lib
from conftest import ROOT, WORLDS, has_cuda, synth_world, world_text


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "prl.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(prl_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = C.CDLL(_lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), f"{n} is declared in include/prl.h but not exported by libprl.so"
    assert set(names) == set(_lib.EXPORTED_SYMBOLS)
    assert _lib.lib().prl_abi_version() == 2


def test_rust_facade_binds_exactly_the_exported_symbols():
    # rust/src/ffi.rs is the extern "C" block a maintainer adds to the reference crate
    text = open(os.path.join(ROOT, "rust", "src", "ffi.rs")).read()
    rust = sorted(set(re.findall(r"pub fn (prl_[a-z0-9_]+)\(", text)))
    assert rust == declared_symbols()


def test_pod_layouts_agree_between_c_rust_and_ctypes():
    """prl_object / prl_ga_config cross the ABI by pointer: the C header (static_assert in
    csrc/prl_world.cc), rust/src/ffi.rs (const assertions) and the ctypes mirror must agree."""
    text = open(os.path.join(ROOT, "rust", "src", "ffi.rs")).read()
    rust = {k: int(v) for k, v in re.findall(r"pub const (PRL_[A-Z_]+): \w+ = (\d+);", text)}
    assert rust["PRL_OBJECT_SIZE"] == C.sizeof(_lib.PrlObject) == 28
    assert rust["PRL_GA_CONFIG_SIZE"] == C.sizeof(_lib.PrlGaConfig) == 40
    assert rust["PRL_ABI_VERSION"] == _lib.lib().prl_abi_version()
    assert [getattr(_lib.PrlObject, f).offset for f, _ in _lib.PrlObject._fields_] == [0, 12, 20, 24]
    assert [getattr(_lib.PrlGaConfig, f).offset for f, _ in _lib.PrlGaConfig._fields_] == [0, 8, 16, 20, 24, 32]
    cc = open(os.path.join(ROOT, "physics_reinforcement_learning_environment_b200", "csrc", "prl_world.cc")).read()
    assert "static_assert(sizeof(prl_object) == 28" in cc and "static_assert(sizeof(prl_ga_config) == 40" in cc


def test_product_does_not_link_or_import_the_oracle():
    # the oracle is test infrastructure: no product source may mention it
    pkg = os.path.join(ROOT, "physics_reinforcement_learning_environment_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cc", ".h")) or f == "Makefile":
                src = open(os.path.join(dirpath, f)).read()
                assert "prl_oracle" not in src and "import oracle" not in src and "from oracle" not in src, f


@pytest.mark.parametrize("name", WORLDS)
def test_world_json_round_trip_is_byte_exact(name):
    text = world_text(name).strip()
    w = prl.World.from_json(text)
    assert w.to_json() == text  # serde_json::to_string formatting, including the "-0.0" rotations
    assert json.loads(w.to_json()) == json.loads(text)


def test_world_counts_and_handles():
    # Appendix B of SURVEY.md
    assert prl.World.from_json(world_text("right")).counts() == (1, 2, 1, 1)
    assert prl.World.from_json(world_text("jump")).counts() == (1, 5, 1, 4)
    push = prl.World.from_json(world_text("push"))
    assert push.counts() == (2, 7, 1, 11)
    assert push.rigid_body_handles() == [None, 1, None, None, None, None, None]  # common.rs:162-172
    assert prl.World.from_json(json.dumps(synth_world())).counts() == (1, 65, 1, 64)


def test_world_create_matches_from_json():
    d = json.loads(world_text("push"))
    objs = []
    for o in d["objects"]:
        kind = prl.WorldObject.GOAL if o["object"] == "Goal" else prl.WorldObject.block(o["object"]["Block"]["fixed"])
        objs.append(prl.ObjectAndTransform(kind, o["position"], o["scale"], o["rotation"]))
    w = prl.World(d["player_position"], objs)
    assert w.to_json() == world_text("push").strip()
    assert [o.object for o in w.objects] == [0, 1, 0, 0, 0, 0, 2]


OBJ = '{"object":"Goal","position":[1,2,3],"scale":[4,5],"rotation":6}'


@pytest.mark.parametrize("bad", [
    "", "{", "[]", '{"player_position":[0,0]}', '{"player_position":[0,0,0],"objects":[]}',
    '{"player_position":[0,0],"objects":[{"object":"Gaol","position":[0,0,0],"scale":[1,1],"rotation":0}]}',
    '{"player_position":[0,0],"objects":[{"object":{"Block":{"fixd":true}},"position":[0,0,0],"scale":[1,1],"rotation":0}]}',
    '{"player_position":[0,0],"objects":[{"object":"Goal","position":[0,0],"scale":[1,1],"rotation":0}]}',
    # derive(Deserialize): a known field twice is "duplicate field"
    '{"player_position":[0,0],"objects":[],"objects":[]}',
    '{"player_position":[0,0],"objects":[{"object":"Goal","object":"Goal","position":[0,0,0],"scale":[1,1],"rotation":0}]}',
    # serde_json number grammar (RFC 8259): no leading '.', '+' or zeros, no trailing '.', no NaN, f64 overflow
    '{"player_position":[.5,0],"objects":[]}', '{"player_position":[01,0],"objects":[]}', '{"player_position":[1.,0],"objects":[]}',
    '{"player_position":[+1,0],"objects":[]}', '{"player_position":[NaN,0],"objects":[]}', '{"player_position":[1e999,0],"objects":[]}',
    '{"player_position":[-,0],"objects":[]}', '{"player_position":[1e,0],"objects":[]}',
    # enum forms serde rejects: a struct variant as a bare string, a unit variant with a payload, two tags, no tag
    '{"player_position":[0,0],"objects":[{"object":"Block","position":[0,0,0],"scale":[1,1],"rotation":0}]}',
    '{"player_position":[0,0],"objects":[{"object":{"Goal":{}},"position":[0,0,0],"scale":[1,1],"rotation":0}]}',
    '{"player_position":[0,0],"objects":[{"object":{"Goal":null,"Block":{"fixed":true}},"position":[0,0,0],"scale":[1,1],"rotation":0}]}',
    '{"player_position":[0,0],"objects":[{"object":{},"position":[0,0,0],"scale":[1,1],"rotation":0}]}',
    # trailing characters, trailing commas, bad escapes, a lone surrogate, a truncated literal
    '{"player_position":[0,0],"objects":[]} x', '{"player_position":[0,0],"objects":[],}', '{"player_position":[0,0,],"objects":[]}',
    r'{"player_position":[0,0],"objects":[],"a":"\x"}', r'{"player_position":[0,0],"objects":[],"a":"\ud800"}',
    r'{"player_position":[0,0],"objects":[],"a":"\u00"}', '{"player_position":[0,0],"objects":[],"a":tru}',
    '{"player_position":[0,0],"objects":[],"a":' + "[" * 200 + "]" * 200 + '}',  # recursion limit 128
    '[[0,0]]', '[[0,0],[],[]]',  # a struct as a sequence needs exactly its fields
    '{"player_position":[0,0],"objects":[],"a":"line\nbreak"}',  # unescaped control character in a string
    b'{"player_position":[0,0],"objects":[],"a":"\xff"}',  # not UTF-8 (from_str takes a &str)
])
def test_malformed_json_is_a_parse_error(bad):
    with pytest.raises(prl.PrlError) as e:
        prl.World.from_json(bad)
    assert e.value.status == _lib.PRL_ERR_PARSE


@pytest.mark.parametrize("text,n_objects", [
    # unknown keys are ignored at every level (no deny_unknown_fields on common.rs:19-49), whatever their value
    ('{"player_position":[0,0],"objects":[],"extra":1}', 0),
    (r'{"version":{"major":[1,2,{"x":null}],"s":"é😀 \"q\" \n"},"player_position":[0,0],"objects":[%s]}' % OBJ, 1),
    ('{"player_position":[0,0],"objects":[{"object":"Goal","position":[1,2,3],"colour":"red","scale":[4,5],"rotation":6,"rotation_deg":[]}]}', 1),
    ('{"player_position":[0,0],"objects":[{"object":{"Block":{"fixed":true,"material":{"friction":0.5}}},"position":[1,2,3],"scale":[4,5],"rotation":6}]}', 1),
    # unknown keys may repeat
    ('{"x":1,"x":2,"player_position":[0,0],"objects":[]}', 0),
    # the unit variant in map form; the struct variant and the structs as sequences (serde visit_seq)
    ('{"player_position":[0,0],"objects":[{"object":{"Goal":null},"position":[1,2,3],"scale":[4,5],"rotation":6}]}', 1),
    ('{"player_position":[0,0],"objects":[{"object":{"Block":[false]},"position":[1,2,3],"scale":[4,5],"rotation":6}]}', 1),
    ('[[0,0],[["Goal",[1,2,3],[4,5],6]]]', 1),
    # escaped keys compare after unescaping; integers, exponents and -0 are numbers; whitespace anywhere
    (r'{"play\u0065r_position":[0,0],"objects":[]}', 0),
    (' {\t"player_position" : [ 1e2 , -0 ] ,\r\n "objects" : [ ] } \n', 0),
    ('{"player_position":[1E+2,1e-2],"objects":[]}', 0),
    ('{"player_position":[123456789012345678901234567890,1e300],"objects":[]}', 0),  # f64 -> `as f32` (inf), accepted
    ('{"player_position":[0,0],"objects":[],"a":' + "[" * 100 + "]" * 100 + '}', 0),
])
def test_json_ingest_follows_serde(text, n_objects):
    w = prl.World.from_json(text)
    assert len(w.objects) == n_objects
    if n_objects:
        o = w.objects[0]
        assert list(o.position) == [1.0, 2.0, 3.0] and list(o.scale) == [4.0, 5.0] and o.rotation == 6.0


def test_capacity_limits_are_reported():
    blocks = [{"object": {"Block": {"fixed": False}}, "position": [60.0 * i, 0.0, 1.0], "scale": [20.0, 20.0], "rotation": 0.0}
              for i in range(8)]
    with pytest.raises(prl.PrlError) as e:
        prl.World.from_json(json.dumps({"player_position": [0, 0], "objects": blocks}))
    assert e.value.status == _lib.PRL_ERR_CAPACITY


def test_move_bits():
    assert prl.Move().bits() == 0 and prl.Move(left=True, up=True).bits() == 5
    assert prl.Move.from_bits(6) == prl.Move(False, True, True)


@pytest.mark.skipif(has_cuda(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    w = prl.World.from_json(world_text("right"))
    with pytest.raises(prl.PrlError) as e:
        prl.BatchedEnvironment(w, 8)
    assert e.value.status == _lib.PRL_ERR_CUDA


def ryu_f32(x):
    """serde_json / ryu::pretty::format32 restated on numpy's shortest round-trip digits."""
    x = np.float32(x)
    if not np.isfinite(x):
        return "null"
    sign = "-" if np.signbit(x) else ""
    if x == 0:
        return sign + "0.0"
    sci = np.format_float_scientific(abs(x), unique=True, trim="-")   # e.g. '1.234e-07', '1e+30'
    mant, exp = sci.split("e")
    digits = mant.replace(".", "")
    kk = int(exp) + 1
    k = kk - len(digits)
    if 0 <= k and kk <= 13:
        return sign + digits + "0" * k + ".0"
    if 0 < kk <= 13:
        return sign + digits[:kk] + "." + digits[kk:]
    if -6 < kk <= 0:
        return sign + "0." + "0" * (-kk) + digits
    return sign + digits[0] + ("." + digits[1:] if len(digits) > 1 else "") + "e" + str(kk - 1)


def test_f32_formatting_follows_ryu():
    """serde_json::to_string writes f32 with ryu: known spellings, the layout thresholds (13 integer digits,
    five leading zeros), and 20000 random bit patterns that must also parse back to the same bits."""
    known = {1.0: "1.0", -0.0: "-0.0", 0.5: "0.5", 1e12: "1000000000000.0", 1e13: "1e13", 123456.79: "123456.79",
             1e-5: "0.00001", 1e-6: "0.000001", 1e-7: "1e-7", 1.234e-7: "1.234e-7", 3.4028235e38: "3.4028235e38",
             1e-45: "1e-45", 16777216.0: "16777216.0", 0.1: "0.1", 1234.5: "1234.5"}
    rng = np.random.Generator(np.random.Philox(2024))
    bits = rng.integers(0, 2**32, size=20000, dtype=np.uint64).astype(np.uint32)
    values = [np.float32(v) for v in known] + [b for b in bits.view(np.float32) if np.isfinite(b)]
    for lo in range(0, len(values), 120):
        chunk = values[lo:lo + 120]
        # position[2] is the editor's draw order: stored and written back, ignored by the physics
        objs = [prl.ObjectAndTransform(prl.WorldObject.block(True), [0.0, 0.0, float(v)], [1.0, 1.0], 0.0) for v in chunk]
        text = prl.World([0.0, 0.0], objs).to_json()
        got = [o["position"][2] for o in json.loads(text, parse_float=lambda t: t, parse_int=lambda t: t)["objects"]]
        for v, g in zip(chunk, got):
            assert g == ryu_f32(v), (v, g, ryu_f32(v))
        back = prl.World.from_json(text)
        assert [np.float32(o.position[2]).view(np.uint32) for o in back.objects] == [np.float32(v).view(np.uint32) for v in chunk]
    for v, spelled in known.items():
        assert ryu_f32(v) == spelled, (v, ryu_f32(v), spelled)
