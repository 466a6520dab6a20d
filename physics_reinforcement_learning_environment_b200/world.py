"""`World`, `ObjectAndTransform`, `WorldObject`, `Move`: the level description and the action, with
the names, fields and JSON schema of the reference (src/common.rs:19-49, 351-356)."""
import ctypes as C
import dataclasses
import json
from typing import List, Optional, Sequence, Union

from . import _lib


@dataclasses.dataclass(frozen=True)
class Move:
    """common.rs:351-356 -- `Default` is all false. One byte on the wire: bit0 left, bit1 right, bit2 up."""
    left: bool = False
    right: bool = False
    up: bool = False

    def bits(self) -> int:
        return (1 if self.left else 0) | (2 if self.right else 0) | (4 if self.up else 0)

    @staticmethod
    def from_bits(b: int) -> "Move":
        return Move(bool(b & 1), bool(b & 2), bool(b & 4))


class WorldObject:
    """common.rs:45-49: `Block { fixed }` | `Goal` (serde externally tagged)."""
    BLOCK_FIXED = 0
    BLOCK_PUSHABLE = 1
    GOAL = 2

    @staticmethod
    def block(fixed: bool) -> int:
        return WorldObject.BLOCK_FIXED if fixed else WorldObject.BLOCK_PUSHABLE


@dataclasses.dataclass
class ObjectAndTransform:
    """common.rs:26-31"""
    object: int
    position: Sequence[float]
    scale: Sequence[float]
    rotation: float


class World:
    """common.rs:19-23. Owns a `prl_world` handle (the parsed level plus its physics-space tables)."""

    def __init__(self, player_position: Sequence[float] = (0.0, 0.0),
                 objects: Optional[List[ObjectAndTransform]] = None, _handle=None):
        L = _lib.lib()
        if _handle is not None:
            self._h = _handle
            return
        objects = objects or []
        arr = (_lib.PrlObject * max(1, len(objects)))()
        for i, o in enumerate(objects):
            arr[i].position[:] = list(o.position)
            arr[i].scale[:] = list(o.scale)
            arr[i].rotation = o.rotation
            arr[i].kind = int(o.object)
        pp = (C.c_float * 2)(*player_position)
        h = C.c_void_p()
        _lib.check(L.prl_world_create(pp, arr, len(objects), C.byref(h)))
        self._h = h

    @staticmethod
    def from_json(text: Union[str, bytes]) -> "World":
        """serde_json::from_str::<World> (editor.rs:916-936)."""
        data = text.encode("utf-8") if isinstance(text, str) else bytes(text)
        h = C.c_void_p()
        _lib.check(_lib.lib().prl_world_from_json(data, len(data), C.byref(h)))
        return World(_handle=h)

    @staticmethod
    def load(path: str) -> "World":
        with open(path, "rb") as f:
            return World.from_json(f.read())

    def to_json(self) -> str:
        """serde_json::to_string(&World) (editor.rs:938-962)."""
        L = _lib.lib()
        need = C.c_size_t()
        _lib.check(L.prl_world_to_json(self._h, None, 0, C.byref(need)))
        buf = C.create_string_buffer(need.value)
        _lib.check(L.prl_world_to_json(self._h, buf, need.value, C.byref(need)))
        return buf.value.decode("utf-8")

    def to_dict(self) -> dict:
        return json.loads(self.to_json())

    @property
    def player_position(self):
        out = (C.c_float * 2)()
        _lib.check(_lib.lib().prl_world_player_position(self._h, out))
        return [out[0], out[1]]

    @property
    def objects(self) -> List[ObjectAndTransform]:
        L = _lib.lib()
        res = []
        for i in range(L.prl_world_num_objects(self._h)):
            o = _lib.PrlObject()
            _lib.check(L.prl_world_get_object(self._h, i, C.byref(o)))
            res.append(ObjectAndTransform(o.kind, list(o.position), list(o.scale), o.rotation))
        return res

    def counts(self):
        """(bodies, colliders, goals, candidate pairs) of Environment::from_world(self)."""
        v = [C.c_int32() for _ in range(4)]
        _lib.check(_lib.lib().prl_world_counts(self._h, *[C.byref(x) for x in v]))
        return tuple(x.value for x in v)

    def rigid_body_handles(self) -> List[Optional[int]]:
        """from_world's `Vec<Option<RigidBodyHandle>>` (common.rs:162-172) as body indices."""
        L = _lib.lib()
        out = []
        for i in range(L.prl_world_num_objects(self._h)):
            b = L.prl_world_object_body(self._h, i)
            out.append(None if b < 0 else b)
        return out

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                _lib.lib().prl_world_destroy(h)
            except Exception:
                pass
            self._h = None
