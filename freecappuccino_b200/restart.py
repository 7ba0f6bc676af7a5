"""Restart files in the reference's own format: ``write_restart_files`` / ``readfiles``
(src/write_restart_files.f90:16-62, src/readfiles.f90:12-50) write one Fortran *unformatted sequential* record per
array -- with gfortran a 4-byte little-endian byte count before and after the payload -- in the fixed order

    (itime, time) [gradpcmf if const_mflux] flmass u v w p te ed t vis uu vv ww uv uw vw uo vo wo teo edo

Arrays the run has not allocated (``t`` without the energy equation, the Reynolds stresses in laminar runs) appear as
empty records.  Host-side I/O only; nothing here touches the GPU path."""
from __future__ import annotations

import struct
from typing import Dict, Optional

import numpy as np

ORDER = ("flmass", "u", "v", "w", "p", "te", "ed", "t", "vis", "uu", "vv", "ww", "uv", "uw", "vw", "uo", "vo", "wo",
         "teo", "edo")
_MAX_RECORD = 2 ** 31 - 9      # longer records are split into sub-records by gfortran; not needed below 268 M values


def _record(payload: bytes) -> bytes:
    if len(payload) > _MAX_RECORD:
        raise ValueError("record longer than 2 GiB: gfortran sub-records are not implemented")
    n = struct.pack("<i", len(payload))
    return n + payload + n


def write_restart(path: str, itime: int, time: float, fields: Dict[str, Optional[np.ndarray]], const_mflux: bool = False,
                  gradpcmf: float = 0.0) -> None:
    """``fields``: name -> float64 array (or None / missing for an unallocated array) for the names in ORDER."""
    with open(path, "wb") as fh:
        fh.write(_record(struct.pack("<id", int(itime), float(time))))
        if const_mflux:
            fh.write(_record(struct.pack("<d", float(gradpcmf))))
        for name in ORDER:
            a = fields.get(name)
            fh.write(_record(b"" if a is None else np.ascontiguousarray(a, dtype="<f8").tobytes()))


def read_restart(path: str, const_mflux: bool = False) -> Dict[str, object]:
    out: Dict[str, object] = {}
    with open(path, "rb") as fh:
        def rec() -> bytes:
            head = fh.read(4)
            if len(head) != 4:
                raise ValueError("unexpected end of restart file")
            (n,) = struct.unpack("<i", head)
            if n < 0:
                raise ValueError("gfortran sub-records are not implemented")
            payload = fh.read(n)
            (m,) = struct.unpack("<i", fh.read(4))
            if m != n or len(payload) != n:
                raise ValueError("corrupt record markers")
            return payload
        out["itime"], out["time"] = struct.unpack("<id", rec())
        if const_mflux:
            (out["gradpcmf"],) = struct.unpack("<d", rec())
        for name in ORDER:
            out[name] = np.frombuffer(rec(), dtype="<f8").copy()
        if fh.read(1):
            raise ValueError("trailing data after the last record")
    return out
