"""Minimal reader for R `save()` files (.rda), enough to pull numeric matrices, factors, lists and
S4 slots out of the reference's bundled fixtures without R.

Test infrastructure only: used by `make_golden.py` in the build container, where the reference
checkout is mounted.  Nothing on the product path imports this.

Format (R Internals, "Serialization Formats"): gzip/xz/bzip2 stream -> "RDX2\\n"/"RDX3\\n" magic ->
"X\\n" (XDR, big-endian) -> three int32 versions (+ native encoding string for v3) -> one pairlist of
(tag symbol -> value).  Every item starts with an int32 of flags: type = low 8 bits, bit 8 =
is-object, bit 9 = has-attributes, bit 10 = has-tag.
"""
from __future__ import annotations

import bz2
import gzip
import lzma
import struct

import numpy as np

NILVALUE_SXP = 254
REFSXP = 255
GLOBALENV_SXP = 253
EMPTYENV_SXP = 242
BASEENV_SXP = 241
MISSINGARG_SXP = 251
UNBOUNDVALUE_SXP = 252
NAMESPACESXP = 249
PACKAGESXP = 250
PERSISTSXP = 247
ALTREP_SXP = 238
ATTRLISTSXP = 239
ATTRLANGSXP = 240
BASENAMESPACE_SXP = 244

NA_INT = -2147483648


class RObj:
    """A decoded R value: `.value` plus `.attr` (dict of attribute name -> RObj/py value)."""

    __slots__ = ("value", "attr", "rtype")

    def __init__(self, value, attr=None, rtype=None):
        self.value = value
        self.attr = attr or {}
        self.rtype = rtype

    def __repr__(self):
        v = self.value
        if isinstance(v, np.ndarray):
            v = f"ndarray{v.shape}:{v.dtype}"
        return f"RObj(type={self.rtype}, value={v}, attr={list(self.attr)})"


class _Reader:
    def __init__(self, buf: bytes):
        self.b = buf
        self.p = 0
        self.refs = []

    def i32(self) -> int:
        (v,) = struct.unpack_from(">i", self.b, self.p)
        self.p += 4
        return v

    def length(self) -> int:
        n = self.i32()
        if n == -1:
            hi = self.i32()
            lo = self.i32()
            n = (hi << 32) + (lo & 0xFFFFFFFF)
        return n

    def raw(self, n: int) -> bytes:
        v = self.b[self.p:self.p + n]
        self.p += n
        return v

    def item(self):
        flags = self.i32()
        t = flags & 0xFF
        is_obj = bool(flags & (1 << 8))
        has_attr = bool(flags & (1 << 9))
        has_tag = bool(flags & (1 << 10))
        del is_obj

        if t == NILVALUE_SXP:
            return None
        if t in (GLOBALENV_SXP, EMPTYENV_SXP, BASEENV_SXP, MISSINGARG_SXP, UNBOUNDVALUE_SXP,
                 BASENAMESPACE_SXP):
            return RObj(None, rtype=t)
        if t == REFSXP:
            idx = flags >> 8
            if idx == 0:
                idx = self.i32()
            return self.refs[idx - 1]
        if t in (NAMESPACESXP, PACKAGESXP, PERSISTSXP):
            # STRSXP-like info block: int 0, length, then CHARSXP items
            self.i32()
            n = self.i32()
            info = [self.item() for _ in range(n)]
            obj = RObj(info, rtype=t)
            self.refs.append(obj)
            return obj
        if t == 1:  # SYMSXP
            name = self.item()
            obj = RObj(name, rtype=1)
            self.refs.append(obj)
            return obj
        if t in (2, 6, 5, ATTRLISTSXP, ATTRLANGSXP):  # pairlist-like / language / promise
            attr = {}
            if t in (ATTRLISTSXP, ATTRLANGSXP):
                has_attr = True
            if has_attr:
                attr = self._attrs(self.item())
            tag = None
            if has_tag:
                tag = self.item()
            car = self.item()
            cdr = self.item()
            items = [(tag.value if isinstance(tag, RObj) else tag, car)]
            if isinstance(cdr, RObj) and cdr.rtype == 2:
                items.extend(cdr.value)
            return RObj(items, attr, rtype=2)
        if t == 4:  # ENVSXP
            self.i32()  # locked
            obj = RObj({}, rtype=4)
            self.refs.append(obj)
            enclos = self.item()
            frame = self.item()
            hashtab = self.item()
            attrib = self.item()
            del enclos, hashtab, attrib
            if isinstance(frame, RObj) and frame.rtype == 2:
                obj.value = {k: v for k, v in frame.value}
            return obj
        if t == 9:  # CHARSXP
            n = self.i32()
            if n == -1:
                return None  # NA_character_
            return self.raw(n).decode("utf-8", "replace")
        if t == 10 or t == 13:  # LGLSXP / INTSXP
            n = self.length()
            v = np.frombuffer(self.b, dtype=">i4", count=n, offset=self.p).astype(np.int32)
            self.p += 4 * n
            return self._finish(v, has_attr, t)
        if t == 14:  # REALSXP
            n = self.length()
            v = np.frombuffer(self.b, dtype=">f8", count=n, offset=self.p).astype(np.float64)
            self.p += 8 * n
            return self._finish(v, has_attr, t)
        if t == 16:  # STRSXP
            n = self.length()
            v = [self.item() for _ in range(n)]
            return self._finish(v, has_attr, t)
        if t == 19 or t == 20:  # VECSXP / EXPRSXP
            n = self.length()
            v = [self.item() for _ in range(n)]
            return self._finish(v, has_attr, t)
        if t == 24:  # RAWSXP
            n = self.length()
            v = self.raw(n)
            return self._finish(v, has_attr, t)
        if t == 25:  # S4SXP: slots live in the attributes
            return self._finish(None, has_attr, t)
        if t == ALTREP_SXP:
            info = self.item()
            state = self.item()
            attr = self.item()
            cls = info.value[0][1].value if isinstance(info, RObj) else None
            val = self._altrep(cls, state)
            if isinstance(attr, RObj) and attr.rtype == 2:
                val.attr.update(self._attrs(attr))
            return val
        if t == 3 or t == 7 or t == 8:  # closures / builtins: not needed, but keep the stream aligned
            if t == 3:
                attr = self.item() if has_attr else None
                env = self.item()
                formals = self.item()
                body = self.item()
                del attr, env, formals, body
                return RObj(None, rtype=3)
            n = self.i32()
            return RObj(self.raw(n).decode(), rtype=t)
        if t == 21:  # BCODESXP
            raise NotImplementedError("byte code in .rda not supported")
        raise NotImplementedError(f"SEXP type {t} at offset {self.p}")

    def _altrep(self, cls, state):
        if cls in ("compact_intseq", "compact_realseq"):
            n, start, step = (float(x) for x in state.value[:3])
            v = start + step * np.arange(int(n))
            if cls == "compact_intseq":
                return RObj(v.astype(np.int32), rtype=13)
            return RObj(v.astype(np.float64), rtype=14)
        if cls is not None and cls.startswith("wrap_"):
            inner = state.value[0] if isinstance(state.value, list) and not isinstance(state.value[0], tuple) \
                else state.value[0][1]
            return RObj(inner.value, dict(inner.attr), inner.rtype)
        if cls == "deferred_string":
            inner = state.value[0][1]
            return RObj([str(x) for x in inner.value], rtype=16)
        raise NotImplementedError(f"ALTREP class {cls}")

    def _attrs(self, pl):
        out = {}
        if isinstance(pl, RObj) and pl.rtype == 2:
            for k, v in pl.value:
                out[k] = v
        return out

    def _finish(self, v, has_attr, t):
        attr = {}
        if has_attr:
            attr = self._attrs(self.item())
        return RObj(v, attr, rtype=t)


def _decompress(raw: bytes) -> bytes:
    if raw[:2] == b"\x1f\x8b":
        return gzip.decompress(raw)
    if raw[:6] == b"\xfd7zXZ\x00":
        return lzma.decompress(raw)
    if raw[:3] == b"BZh":
        return bz2.decompress(raw)
    return raw


def read_rda(path: str) -> dict:
    """Return {object name: RObj} for every object saved in `path`."""
    buf = _decompress(open(path, "rb").read())
    if buf[:5] not in (b"RDX2\n", b"RDX3\n"):
        raise ValueError(f"{path}: not an RDX2/RDX3 file")
    r = _Reader(buf)
    r.p = 5
    fmt = r.raw(2)
    if fmt != b"X\n":
        raise ValueError("only XDR serialisation is supported")
    version = r.i32()
    r.i32()  # writer R version
    r.i32()  # min reader version
    if version == 3:
        n = r.i32()
        r.raw(n)
    top = r.item()
    return {k: v for k, v in top.value}


def as_matrix(obj: RObj) -> np.ndarray:
    """R column-major vector with a `dim` attribute -> numpy (rows, cols) array (Fortran order kept)."""
    dim = obj.attr["dim"].value
    return np.asarray(obj.value).reshape(tuple(int(d) for d in dim), order="F")


def dimnames(obj: RObj):
    dn = obj.attr.get("dimnames")
    if dn is None:
        return None, None
    rows, cols = dn.value
    return (rows.value if rows is not None else None), (cols.value if cols is not None else None)


def as_factor(obj: RObj):
    """Factor -> (codes 1-based int array, levels list)."""
    return np.asarray(obj.value), obj.attr["levels"].value


def named_list(obj: RObj) -> dict:
    names = obj.attr["names"].value
    return {n: v for n, v in zip(names, obj.value)}
