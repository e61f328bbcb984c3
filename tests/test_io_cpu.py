"""Host-side ingest / checkpoint, no device needed:
* shud_b200_format_ic writes the reference's initial-condition file byte for byte (the fixture holds the file the
  reference's own Model_Data::PrintInit wrote for the same state, via oracle/ref_driver.cpp --print-init);
* the binary mesh container round-trips every array of shud_mesh and rejects damaged files."""
import os
import time

import numpy as np
import pytest

from shud_up_b200 import abi, api, synth

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_ic_file_is_byte_identical_to_the_reference_writer(tmp_path):
    g = np.load(os.path.join(GOLD, "qhh.icfile.npz"))
    mesh = np.load(os.path.join(GOLD, "qhh.mesh.npz"))
    Ne, Nr, Nl = int(mesh["Ne"][0]), int(mesh["Nr"][0]), int(mesh["Nl"][0])
    out = tmp_path / "qhh.cfg.ic.update"
    api.format_ic(out, float(g["ic_t"][0]), Ne, Nr, Nl, g["y"], g["ic_yEleIS"], g["ic_yEleSnow"])
    assert open(out, "rb").read() == g["text"].tobytes()


def test_ic_file_reads_back(tmp_path):
    """restart: the reference-written file parses to the printed (6-decimal) values; a file of another mesh is refused"""
    g = np.load(os.path.join(GOLD, "qhh.icfile.npz"))
    mesh = np.load(os.path.join(GOLD, "qhh.mesh.npz"))
    Ne, Nr, Nl = int(mesh["Ne"][0]), int(mesh["Nr"][0]), int(mesh["Nl"][0])
    path = tmp_path / "qhh.cfg.ic"
    open(path, "wb").write(g["text"].tobytes())
    t, y, ics, snow = api.read_ic(path, Ne, Nr, Nl)
    assert t == float(g["ic_t"][0])
    assert np.allclose(y, g["y"], rtol=0, atol=5.0000001e-7) and np.allclose(ics, g["ic_yEleIS"], rtol=0, atol=5.0000001e-7)
    assert np.allclose(snow, g["ic_yEleSnow"], rtol=0, atol=5.0000001e-7)
    # writing what was read reproduces the file byte for byte (the text is a fixed point of read -> format)
    again = tmp_path / "again.ic"
    api.format_ic(again, t, Ne, Nr, Nl, y, ics, snow)
    assert open(again, "rb").read() == g["text"].tobytes()
    with pytest.raises(RuntimeError):
        api.read_ic(path, Ne - 1, Nr, Nl)


def _fields(mesh):
    m, keep = abi.make_mesh(mesh)
    Ne, Nr, Ns, Nl = m.Ne, m.Nr, m.Ns, m.Nl
    nb = int(np.asarray(mesh["lake_bathy_ptr"])[Nl]) if Nl else 0
    dims = {}
    for n in abi.MESH_CELL_D + abi.MESH_CELL_I + ["x", "y"]:
        dims[n] = Ne
    for n in abi.MESH_EDGE_D + abi.MESH_EDGE_I:
        dims[n] = 3 * Ne
    for n in abi.MESH_RIV_D + abi.MESH_RIV_I:
        dims[n] = Nr
    for n in abi.MESH_SEG_D + abi.MESH_SEG_I:
        dims[n] = Ns
    dims.update(lake_zmin=Nl, lake_NumEleLake=Nl, lake_bathy_ptr=Nl + 1 if Nl else 0, lake_bathy_yi=nb, lake_bathy_ai=nb)
    return m, keep, dims


@pytest.mark.parametrize("basin", ["ccw", "qhh"])
def test_mesh_container_round_trip(tmp_path, basin):
    mesh = dict(np.load(os.path.join(GOLD, f"{basin}.mesh.npz")))
    m, keep, dims = _fields(mesh)
    path = tmp_path / f"{basin}.shudb200"
    api.mesh_save(path, mesh)
    ld = api.LoadedMesh(path)
    try:
        for n in ("Ne", "Nr", "Ns", "Nl", "close_boundary", "lakeon"):
            assert getattr(ld.mesh, n) == getattr(m, n), n
        for n, cnt in dims.items():
            got = ld.array(n, cnt)
            src = getattr(m, n)
            if not src or cnt == 0:
                assert got is None or cnt == 0, n
                continue
            ref = np.ctypeslib.as_array(src, shape=(cnt,))
            assert np.array_equal(got, ref), n
    finally:
        ld.close()
    # damaged files are refused
    raw = open(path, "rb").read()
    bad = tmp_path / "bad.shudb200"
    open(bad, "wb").write(raw[: len(raw) // 2])
    with pytest.raises(RuntimeError):
        api.LoadedMesh(bad)
    open(bad, "wb").write(b"NOTSHUD!" + raw[8:])
    with pytest.raises(RuntimeError):
        api.LoadedMesh(bad)
    # crafted headers: an array offset near 2^64 (would wrap in offset + size), a misaligned one, and a payload size
    # beyond what the file holds must all be refused, not dereferenced
    import struct
    hdr = 8 + 4 + 4 + 6 * 4 + 8 + 8                      # Header: magic, version, nfields, 6 ints, nbathy, payload_bytes
    ent0 = hdr + 24 + 4 + 4 + 8                          # first Entry's `offset` field
    for off in (2 ** 64 - 64, 8):
        b = bytearray(raw); b[ent0:ent0 + 8] = struct.pack("<Q", off)
        open(bad, "wb").write(bytes(b))
        with pytest.raises(RuntimeError):
            api.LoadedMesh(bad)
    b = bytearray(raw); b[hdr - 8:hdr] = struct.pack("<Q", 2 ** 40)
    open(bad, "wb").write(bytes(b))
    with pytest.raises(RuntimeError):
        api.LoadedMesh(bad)


def test_mesh_container_ingest_rate(tmp_path):
    """ingest at file-system speed (informational): a 200k-cell synthetic mesh"""
    mesh = synth.make(400, 250, ntree=10, reaches_per_tree=200)
    path = tmp_path / "m.shudb200"
    api.mesh_save(path, mesh)
    nbytes = os.path.getsize(path)
    t0 = time.perf_counter()
    ld = api.LoadedMesh(path)
    dt = time.perf_counter() - t0
    assert ld.mesh.Ne == int(mesh["Ne"][0])
    ld.close()
    print(f"mesh container: {nbytes / 1e6:.1f} MB loaded in {dt * 1e3:.1f} ms = {nbytes / dt / 1e9:.2f} GB/s")
