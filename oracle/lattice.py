"""Velocity sets and face metadata (oracle).

Tables follow reference ``vivsim/lbm/lattice.py:39-110`` (D2Q9) and
``vivsim/lbm3d/lattice.py:39-126`` (D3Q19).  The per-face direction groups are
*derived* from the velocity table here and checked against the reference's
literal lists by ``tests/test_oracle_golden.py``.
"""

from dataclasses import dataclass, field

import numpy as np


@dataclass(frozen=True)
class Face:
    name: str
    axis: int          # spatial axis normal to the face
    sign: int          # +1: inward normal points along +axis (low side), -1: high side
    wall: int          # index of wall layer along axis (0 or -1)
    neighbor: int      # index of first fluid layer (1 or -2)
    in_dirs: tuple     # populations entering the fluid
    out_dirs: tuple    # populations paired with in_dirs (see below)
    zero_dirs: tuple   # populations with no normal component
    tan_dirs: tuple = ()        # 2-D only: (+t, -t) axis directions
    pos_side_dirs: tuple = ()   # 2-D only: c_t > 0
    neg_side_dirs: tuple = ()   # 2-D only: c_t < 0
    tangential_axes: tuple = ()


@dataclass(frozen=True)
class Lattice:
    d: int
    q: int
    c: np.ndarray
    w: np.ndarray
    opp: np.ndarray
    faces: dict = field(default_factory=dict)

    def face(self, loc):
        return self.faces[loc]   # KeyError for a bad loc, like the reference


def _find(c, vec):
    return int(np.where((c == np.asarray(vec)).all(axis=1))[0][0])


def _opp(c):
    return np.array([_find(c, -v) for v in c], dtype=np.int32)


def _faces_2d(c, opp):
    faces = {}
    for name, axis, sign in (("left", 0, 1), ("right", 0, -1), ("bottom", 1, 1), ("top", 1, -1)):
        t = 1 - axis
        n = np.zeros(2, int); n[axis] = sign
        e = np.zeros(2, int); e[t] = 1
        # in0 = pure normal; in1 carries tangential momentum sign*(+t); in2 the other
        ins = (_find(c, n), _find(c, n + sign * e), _find(c, n - sign * e))
        outs = tuple(int(opp[i]) for i in ins)
        faces[name] = Face(
            name=name, axis=axis, sign=sign,
            wall=0 if sign > 0 else -1, neighbor=1 if sign > 0 else -2,
            in_dirs=ins, out_dirs=outs,
            zero_dirs=(0, _find(c, e), _find(c, -e)),
            tan_dirs=(_find(c, e), _find(c, -e)),
            pos_side_dirs=tuple(int(i) for i in np.where(c[:, t] > 0)[0]),
            neg_side_dirs=tuple(int(i) for i in np.where(c[:, t] < 0)[0]),
            tangential_axes=(t,),
        )
    return faces


def _faces_3d(c):
    faces = {}
    for name, axis, sign in (("left", 0, 1), ("right", 0, -1), ("bottom", 1, 1),
                             ("top", 1, -1), ("back", 2, 1), ("front", 2, -1)):
        ins = tuple(int(i) for i in np.where(c[:, axis] * sign > 0)[0])
        # the reference pairs in_k with its MIRROR image across the wall (normal
        # component flipped, tangential kept), not with the opposite direction
        outs = []
        for i in ins:
            m = c[i].copy(); m[axis] = -m[axis]
            outs.append(_find(c, m))
        faces[name] = Face(
            name=name, axis=axis, sign=sign,
            wall=0 if sign > 0 else -1, neighbor=1 if sign > 0 else -2,
            in_dirs=ins, out_dirs=tuple(outs),
            zero_dirs=tuple(int(i) for i in np.where(c[:, axis] == 0)[0]),
            tangential_axes=tuple(a for a in range(3) if a != axis),
        )
    return faces


_C2 = np.array([[0, 0], [1, 0], [0, 1], [-1, 0], [0, -1],
                [1, 1], [-1, 1], [-1, -1], [1, -1]], dtype=np.int32)
_W2 = np.array([4 / 9] + [1 / 9] * 4 + [1 / 36] * 4, dtype=np.float32)

_C3 = np.array([[0, 0, 0],
                [1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1],
                [1, 1, 0], [-1, 1, 0], [1, -1, 0], [-1, -1, 0],
                [1, 0, 1], [-1, 0, 1], [1, 0, -1], [-1, 0, -1],
                [0, 1, 1], [0, -1, 1], [0, 1, -1], [0, -1, -1]], dtype=np.int32)
_W3 = np.array([1 / 3] + [1 / 18] * 6 + [1 / 36] * 12, dtype=np.float32)

_O2 = _opp(_C2)
D2Q9 = Lattice(d=2, q=9, c=_C2, w=_W2, opp=_O2, faces=_faces_2d(_C2, _O2))
D3Q19 = Lattice(d=3, q=19, c=_C3, w=_W3, opp=_opp(_C3), faces=_faces_3d(_C3))
