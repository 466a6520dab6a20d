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


@pytest.mark.parametrize("bad", [
    "", "{", "[]", '{"player_position":[0,0]}', '{"player_position":[0,0,0],"objects":[]}',
    '{"player_position":[0,0],"objects":[{"object":"Gaol","position":[0,0,0],"scale":[1,1],"rotation":0}]}',
    '{"player_position":[0,0],"objects":[{"object":{"Block":{"fixd":true}},"position":[0,0,0],"scale":[1,1],"rotation":0}]}',
    '{"player_position":[0,0],"objects":[{"object":"Goal","position":[0,0],"scale":[1,1],"rotation":0}]}',
    '{"player_position":[0,0],"objects":[],"extra":1}',
])
def test_malformed_json_is_a_parse_error(bad):
    with pytest.raises(prl.PrlError) as e:
        prl.World.from_json(bad)
    assert e.value.status == _lib.PRL_ERR_PARSE


def test_capacity_limits_are_reported():
    blocks = [{"object": {"Block": {"fixed": False}}, "position": [60.0 * i, 0.0, 1.0], "scale": [20.0, 20.0], "rotation": 0.0}
              for i in range(5)]
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
