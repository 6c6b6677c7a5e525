"""On-disk / wire format for the inputs of the hot path.

The reference has no such format: its inputs only exist as live DOLFINx objects.  One `.eqlb` file holds
everything the C ABI needs to rebuild a problem without DOLFINx - the arrays of `eqlb_mesh`, the degrees that
select `eqlb_tables`, the boundary data, optionally the projected inputs, reference outputs and integer patch
maps - so that a DOLFINx-side user exports a case once (tools/export_dolfinx_fixture.py writes this format) and
parity runs anywhere.

Layout (little endian):
    bytes 0..7    magic  b"EQLBWIRE"
    u32           version (= 1)
    u32           length of the JSON index in bytes
    JSON index    {"meta": {...}, "arrays": {name: {"dtype": "<f8", "shape": [...], "offset": o, "nbytes": n}}}
    padding to a multiple of 64
    raw arrays, each starting at a multiple of 64 (offsets relative to the start of the data section)
Arrays can be memory-mapped (`load(..., mmap=True)`); unknown array names are ignored by readers, new names may be
added without a version bump; a change of the meaning of an existing name bumps the version.

Array names: mesh: x, cell_node, cell_fct, fct_node, fct_cell_off, fct_cell, node_cell_off, node_cell, node_fct_off,
node_fct, fct_perms, cell_perm_info; boundary data: facet_type [nrhs][nfct] i1, bflux_<r> [ncell*nrt] f8 (optional),
local_fct_id [nfct] i1, node_on_stress_bnd [nnode] i1 (optional); data: G_<r>, F_<r>, sigma_<r> (optional);
patch maps (optional, `eqlb_get_patch_maps` layout): pm_ncells, pm_cells, pm_fcts, pm_inodes_local, pm_fcts_local,
pm_type, pm_reversed, pm_reversion.
meta: {"k", "p", "nrhs", "stress", "path": "se"|"ev", "source": free text}
"""

from __future__ import annotations

import json
import struct

import numpy as np

from .mesh import Mesh

MAGIC = b"EQLBWIRE"
VERSION = 1
_ALIGN = 64
MESH_FIELDS = ["x", "cell_node", "cell_fct", "fct_node", "fct_cell_off", "fct_cell", "node_cell_off", "node_cell", "node_fct_off",
               "node_fct", "fct_perms", "cell_perm_info"]


def _pad(n):
    return (-n) % _ALIGN


def save(path, meta: dict, arrays: dict):
    """Write a `.eqlb` file.  `arrays`: name -> numpy array (any of the names above)."""
    index, off = {}, 0
    blobs = []
    for name, a in arrays.items():
        a = np.ascontiguousarray(a)
        if a.dtype.byteorder == ">":
            a = a.astype(a.dtype.newbyteorder("<"))
        index[name] = {"dtype": a.dtype.str.replace("=", "<").replace("|", "|"), "shape": list(a.shape), "offset": off, "nbytes": a.nbytes}
        blobs.append(a)
        off += a.nbytes + _pad(a.nbytes)
    head = json.dumps({"meta": meta, "arrays": index}).encode()
    with open(path, "wb") as fh:
        fh.write(MAGIC)
        fh.write(struct.pack("<II", VERSION, len(head)))
        fh.write(head)
        fh.write(b"\0" * _pad(16 + len(head)))
        for a in blobs:
            fh.write(a.tobytes())
            fh.write(b"\0" * _pad(a.nbytes))


def load(path, mmap=False):
    """Read a `.eqlb` file -> (meta, arrays)."""
    with open(path, "rb") as fh:
        if fh.read(8) != MAGIC:
            raise RuntimeError(f"{path}: not an EQLBWIRE file")
        version, nhead = struct.unpack("<II", fh.read(8))
        if version != VERSION:
            raise RuntimeError(f"{path}: wire format version {version} not supported (reader: {VERSION})")
        head = json.loads(fh.read(nhead).decode())
        base = 16 + nhead + _pad(16 + nhead)
        arrays = {}
        for name, d in head["arrays"].items():
            dt, shape = np.dtype(d["dtype"]), tuple(d["shape"])
            if mmap:
                arrays[name] = np.memmap(path, dtype=dt, mode="r", offset=base + d["offset"], shape=shape)
            else:
                fh.seek(base + d["offset"])
                arrays[name] = np.frombuffer(fh.read(d["nbytes"]), dtype=dt).reshape(shape).copy()
    return head["meta"], arrays


def mesh_arrays(mesh: Mesh) -> dict:
    return {f: getattr(mesh, f) for f in MESH_FIELDS}


def mesh_from_arrays(arrays: dict) -> Mesh:
    """Rebuild a `Mesh` (boundary ids are recomputed from the geometry of the unit-square frame if present)."""
    x = np.asarray(arrays["x"], dtype=np.float64)
    fco = np.asarray(arrays["fct_cell_off"], dtype=np.int32)
    fn = np.asarray(arrays["fct_node"], dtype=np.int32)
    bfct = np.nonzero(np.diff(fco) == 1)[0].astype(np.int32)
    mid = 0.5 * (x[fn[bfct, 0]] + x[fn[bfct, 1]])
    side = np.zeros(bfct.shape[0], dtype=np.int32)
    side[np.isclose(mid[:, 0], 0.0)] = 1
    side[np.isclose(mid[:, 1], 0.0)] = 2
    side[np.isclose(mid[:, 0], 1.0)] = 3
    side[np.isclose(mid[:, 1], 1.0)] = 4
    kw = {f: np.asarray(arrays[f]) for f in MESH_FIELDS}
    return Mesh(bfct=bfct, bfct_side=side, **kw)


def save_problem(path, mesh: Mesh, k, p, bdata, G=None, F=None, sigma=None, patch_maps=None, stress=False, ev=False, source=""):
    """Everything needed to replay one equilibration call."""
    arrays = mesh_arrays(mesh)
    arrays["facet_type"] = np.asarray(bdata.facet_type, dtype=np.int8)
    arrays["local_fct_id"] = np.asarray(bdata.local_fct_id, dtype=np.int8)
    nrhs = arrays["facet_type"].shape[0]
    for r in range(nrhs):
        if bdata.bflux[r] is not None:
            arrays[f"bflux_{r}"] = np.asarray(bdata.bflux[r], dtype=np.float64)
    if getattr(bdata, "node_on_stress_bnd", None) is not None:
        arrays["node_on_stress_bnd"] = np.asarray(bdata.node_on_stress_bnd, dtype=np.int8)
    for tag, vs in (("G", G), ("F", F), ("sigma", sigma)):
        if vs is not None:
            for r, v in enumerate(vs):
                arrays[f"{tag}_{r}"] = np.asarray(v, dtype=np.float64)
    if patch_maps is not None:
        for key in ("ncells", "cells", "fcts", "inodes_local", "fcts_local", "type", "reversed", "reversion"):
            if key in patch_maps:
                arrays["pm_" + key] = np.asarray(patch_maps[key])
    meta = {"k": int(k), "p": int(p), "nrhs": int(nrhs), "stress": bool(stress), "path": "ev" if ev else "se", "source": source}
    save(path, meta, arrays)


class _BD:
    pass


def load_problem(path, mmap=False):
    """-> dict(meta, mesh, bdata, G, F, sigma, patch_maps)"""
    meta, a = load(path, mmap)
    mesh = mesh_from_arrays(a)
    nrhs = meta["nrhs"]
    bd = _BD()
    bd.facet_type = np.ascontiguousarray(a["facet_type"], dtype=np.int8)
    bd.local_fct_id = np.ascontiguousarray(a["local_fct_id"], dtype=np.int8)
    bd.bflux = [np.ascontiguousarray(a[f"bflux_{r}"]) if f"bflux_{r}" in a else None for r in range(nrhs)]
    bd.node_on_stress_bnd = np.ascontiguousarray(a["node_on_stress_bnd"], dtype=np.int8) if "node_on_stress_bnd" in a else None
    bd.num_rhs = nrhs

    def vecs(tag):
        return [np.ascontiguousarray(a[f"{tag}_{r}"]) for r in range(nrhs)] if f"{tag}_0" in a else None

    pm = {key[3:]: np.asarray(v) for key, v in a.items() if key.startswith("pm_")} or None
    return dict(meta=meta, mesh=mesh, bdata=bd, G=vecs("G"), F=vecs("F"), sigma=vecs("sigma"), patch_maps=pm)
