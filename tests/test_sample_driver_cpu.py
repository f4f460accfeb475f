"""Host side of the sampling driver: the NIfTI-1 reader / writer against the published format,
the BraTS case loader's item layout (bratsloader.py:9-109 of the reference) and the case bookkeeping."""
import gzip
import struct

import numpy as np
import pytest
import torch

from fcwdm import nifti


def test_nifti_round_trip_gz_and_plain(tmp_path):
    g = np.random.default_rng(0)
    a = g.random((7, 5, 3)).astype(np.float32)
    for name in ("a.nii.gz", "a.nii"):
        path = tmp_path / name
        nifti.write(path, a)
        b, h = nifti.read(path, return_header=True)
        assert b.dtype == np.float64 and b.shape == a.shape          # get_fdata() semantics
        assert np.array_equal(b, a.astype(np.float64))
        assert h.shape == a.shape and h.datatype == 16 and h.bitpix == 32 and h.vox_offset == 352.0
        assert h.sform_code == 2 and np.array_equal(h.affine, np.eye(4))
    raw = gzip.decompress(open(tmp_path / "a.nii.gz", "rb").read())
    assert raw == open(tmp_path / "a.nii", "rb").read()
    # the bytes follow the nifti_1_header layout: sizeof_hdr, dim[], magic, x fastest in the data block
    assert struct.unpack_from("<i", raw, 0)[0] == 348
    assert struct.unpack_from("<8h", raw, 40) == (3, 7, 5, 3, 1, 1, 1, 1)
    assert raw[344:348] == b"n+1\x00" and len(raw) == 352 + a.size * 4
    first = np.frombuffer(raw, dtype="<f4", count=8, offset=352)
    assert np.array_equal(first[:7], a[:, 0, 0]) and first[7] == a[0, 1, 0]


def test_nifti_reads_big_endian_int16_with_scaling(tmp_path):
    shape = (4, 3, 2)
    vals = np.arange(24, dtype=np.int16).reshape(shape[::-1]).T      # file order: first index fastest
    hdr = bytearray(352)
    struct.pack_into(">i", hdr, 0, 348)
    struct.pack_into(">8h", hdr, 40, 3, *shape, 1, 1, 1, 1)
    struct.pack_into(">2h", hdr, 70, 4, 16)
    struct.pack_into(">8f", hdr, 76, 1, 2, 2, 2, 1, 1, 1, 1)
    struct.pack_into(">3f", hdr, 108, 352.0, 0.5, 10.0)             # vox_offset, scl_slope, scl_inter
    hdr[344:348] = b"n+1\x00"
    path = tmp_path / "be.nii"
    with open(path, "wb") as f:
        f.write(bytes(hdr) + vals.T.astype(">i2").tobytes())
    out, h = nifti.read(path, return_header=True)
    assert h.byteorder == ">" and out.shape == shape
    assert np.array_equal(out, vals.astype(np.float64) * 0.5 + 10.0)
    assert np.array_equal(np.diag(h.affine), [2.0, 2.0, 2.0, 1.0])  # no s/qform: pixdim scaling


def test_nifti_header_copied_from_another_file(tmp_path):
    aff = np.array([[-1.0, 0, 0, 90.0], [0, -1.0, 0, 126.0], [0, 0, 1.0, -72.0], [0, 0, 0, 1.0]])
    src = tmp_path / "src.nii.gz"
    nifti.write(src, np.zeros((4, 4, 4), np.float32), affine=aff)
    _, like = nifti.read(src, return_header=True)
    assert like.sform_code == 2 and like.qform_code == 0 and np.array_equal(like.affine, aff)
    dst = tmp_path / "dst.nii.gz"
    nifti.write(dst, np.ones((6, 5, 4), np.float32), like=like)
    arr, h = nifti.read(dst, return_header=True)
    assert arr.shape == (6, 5, 4) and np.array_equal(h.affine, aff) and h.pixdim[1:4] == like.pixdim[1:4]


def test_nifti_rejects_garbage(tmp_path):
    p = tmp_path / "bad.nii"
    p.write_bytes(b"\0" * 400)
    with pytest.raises(nifti.NiftiError):
        nifti.read(p)
    with pytest.raises(nifti.NiftiError):
        nifti.build_header((40000, 2, 2), np.float32)
    with pytest.raises(nifti.NiftiError):
        nifti.build_header((2, 2), np.complex64)


def _write_case(root, subject, modalities, seed):
    g = np.random.default_rng(seed)
    folder = root / subject
    folder.mkdir(parents=True)
    vols = {}
    for m in modalities:
        v = (g.random((240, 240, 155)) * 1000.0).astype(np.float32)
        v[:20] = 0
        nifti.write(folder / f"{subject}-{m}.nii.gz", v)
        vols[m] = v
    return vols


def test_brats_loader_item_layout(tmp_path):
    from guided_diffusion.bratsloader import BRATSVolumes, clip_and_normalize
    vols = _write_case(tmp_path / "validation", "BraTS-GLI-00001-000", ("t1n", "t2f", "seg"), 0)
    (tmp_path / "validation" / "BraTS-GLI-00001-000" / "notes.txt").write_text("ignored")
    ds = BRATSVolumes(str(tmp_path), mode="eval")
    assert len(ds) == 1 and set(ds.database[0]) == {"t1n", "t2f", "seg"}
    item = ds[0]
    assert item["t1n"].shape == (1, 224, 224, 160) and item["t1n"].dtype == torch.float32
    assert item["t1c"].shape == (1,) and item["t2w"].shape == (1,)
    assert item["missing"] == "t2w"                                  # the LAST absent modality, as in the reference
    assert item["subj"].endswith("BraTS-GLI-00001-000-t1n.nii.gz")
    want = clip_and_normalize(vols["t1n"].astype(np.float64))[8:-8, 8:-8]
    np.testing.assert_allclose(item["t1n"][0, :, :, :155].numpy(), want.astype(np.float32), rtol=0, atol=1e-7)
    assert float(item["t1n"][0, :, :, 155:].abs().max()) == 0.0 and float(item["t1n"].max()) == 1.0
    assert BRATSVolumes(str(tmp_path), mode="train")[0]["subj"] == "dummy_string"
    raw = BRATSVolumes(str(tmp_path), mode="eval", raw=True)[0]
    assert raw["t2f"].shape == (240, 240, 155) and np.array_equal(raw["t2f"].numpy(), vols["t2f"])


def test_driver_bookkeeping():
    from fcwdm.sample_driver import conditions_for, subject_of
    assert conditions_for("t1n") == ("t1c", "t2w", "t2f")            # sample.py:64-89
    assert conditions_for("t2w") == ("t1n", "t1c", "t2f")
    assert conditions_for("t2f") == ("t1n", "t1c", "t2w")
    with pytest.raises(ValueError):
        conditions_for("seg")
    p = "/data/validation/BraTS-GLI-00001-000/BraTS-GLI-00001-000-t1n.nii.gz"
    assert subject_of(p, "validation/") == "BraTS-GLI-00001-000"     # the reference's [:19] rule
    assert subject_of(p) == "BraTS-GLI-00001-000"
    assert subject_of("/x/caseA/caseA-t1n.nii.gz") == "caseA"


# nifti_1_header as published in nifti1.h (field, byte offset, struct format) -- written out here, independently of
# fcwdm.nifti, as the yardstick for the header bytes.  "parity unpinned" for this one row: nibabel is not installed in
# this image, so the golden below is built from the format's published layout plus the values nibabel's
# Nifti1Image(array, affine) + nib.save are documented to set (sform 'aligned' = 2, qform 'unknown' = 0 with the
# quaternion / offsets / pixdim of the affine still filled in, scl_slope 1 / scl_inter 0 for a float image written
# without scaling, vox_offset 352, four zero extension bytes, everything else zero).
NIFTI1_LAYOUT = [
    ("sizeof_hdr", 0, "i"), ("data_type", 4, "10s"), ("db_name", 14, "18s"), ("extents", 32, "i"),
    ("session_error", 36, "h"), ("regular", 38, "c"), ("dim_info", 39, "B"), ("dim", 40, "8h"),
    ("intent_p1", 56, "f"), ("intent_p2", 60, "f"), ("intent_p3", 64, "f"), ("intent_code", 68, "h"),
    ("datatype", 70, "h"), ("bitpix", 72, "h"), ("slice_start", 74, "h"), ("pixdim", 76, "8f"),
    ("vox_offset", 108, "f"), ("scl_slope", 112, "f"), ("scl_inter", 116, "f"), ("slice_end", 120, "h"),
    ("slice_code", 122, "B"), ("xyzt_units", 123, "B"), ("cal_max", 124, "f"), ("cal_min", 128, "f"),
    ("slice_duration", 132, "f"), ("toffset", 136, "f"), ("glmax", 140, "i"), ("glmin", 144, "i"),
    ("descrip", 148, "80s"), ("aux_file", 228, "24s"), ("qform_code", 252, "h"), ("sform_code", 254, "h"),
    ("quatern_b", 256, "f"), ("quatern_c", 260, "f"), ("quatern_d", 264, "f"), ("qoffset_x", 268, "f"),
    ("qoffset_y", 272, "f"), ("qoffset_z", 276, "f"), ("srow_x", 280, "4f"), ("srow_y", 296, "4f"),
    ("srow_z", 312, "4f"), ("intent_name", 328, "16s"), ("magic", 344, "4s"),
]


def _golden_header(shape, affine):
    values = {name: (0,) * int(fmt[:-1] or 1) if fmt[-1] in "hfiB" else (b"",) for name, _, fmt in NIFTI1_LAYOUT}
    values.update(sizeof_hdr=(348,), regular=(b"\0",), dim=(len(shape), *shape, *([1] * (7 - len(shape)))), datatype=(16,),
                  bitpix=(32,), pixdim=(1.0, *np.sqrt((affine[:3, :3] ** 2).sum(0)), 1.0, 1.0, 1.0, 1.0), vox_offset=(352.0,),
                  scl_slope=(1.0,), scl_inter=(0.0,), qform_code=(0,), sform_code=(2,),
                  qoffset_x=(affine[0, 3],), qoffset_y=(affine[1, 3],), qoffset_z=(affine[2, 3],),
                  srow_x=tuple(affine[0]), srow_y=tuple(affine[1]), srow_z=tuple(affine[2]), magic=(b"n+1\0",))
    out = bytearray(352)
    end = 0
    for name, off, fmt in NIFTI1_LAYOUT:
        assert off == end, name                                       # the table itself is gap-free, 348 bytes
        struct.pack_into("<" + fmt, out, off, *values[name])
        end = off + struct.calcsize("<" + fmt)
    assert end == 348
    return bytes(out)


@pytest.mark.parametrize("affine", [np.eye(4), np.diag([1.0, 1.5, 2.0, 1.0]) + np.array([[0, 0, 0, 3.0], [0, 0, 0, -4.0],
                                                                                          [0, 0, 0, 5.5], [0, 0, 0, 0]])])
def test_nifti_header_bytes_against_the_published_layout(tmp_path, affine):
    """What sample.py:141-145 asks nibabel for -- nib.save(nib.Nifti1Image(arr, np.eye(4)), 'x.nii.gz') -- byte for byte:
    the 352 header bytes, then the voxels in Fortran order as little-endian float32."""
    a = np.random.default_rng(3).random((4, 3, 2)).astype(np.float32)
    golden = _golden_header(a.shape, affine) + a.tobytes(order="F")
    assert nifti.encode(a, affine=affine, gz=False) == golden
    path = tmp_path / "x.nii.gz"
    nifti.write(path, a, affine=affine)
    assert gzip.decompress(path.read_bytes()) == golden
    # and the reader against the same table: every field where the published layout puts it
    h = nifti.parse_header(golden)
    fields = {name: struct.unpack_from("<" + fmt, golden, off) for name, off, fmt in NIFTI1_LAYOUT}
    assert h.dim == fields["dim"] and (h.datatype, h.bitpix) == (fields["datatype"][0], fields["bitpix"][0])
    assert h.pixdim == fields["pixdim"] and h.vox_offset == 352.0 and (h.scl_slope, h.scl_inter) == (1.0, 0.0)
    assert (h.qform_code, h.sform_code) == (0, 2) and np.array_equal(h.affine, affine)


def test_nifti_multi_member_gzip_reads_as_one_stream(tmp_path):
    """Volumes above 4 MB are deflated as independent gzip members on a thread pool (RFC 1952 section 2.2: a gzip file is a
    series of members); gzip.decompress / GzipFile -- what nibabel opens .nii.gz with -- and zlib read them as one stream."""
    import zlib
    a = np.random.default_rng(5).random((96, 96, 160)).astype(np.float32)          # 5.9 MB -> two members
    path = tmp_path / "big.nii.gz"
    nifti.write(path, a)
    blob = path.read_bytes()
    plain = nifti.encode(a, gz=False)
    assert gzip.decompress(blob) == plain
    with gzip.GzipFile(path) as f:
        assert f.read() == plain
    members, rest = 0, blob
    out = b""
    while rest:
        d = zlib.decompressobj(16 + zlib.MAX_WBITS)
        out += d.decompress(rest)
        rest = d.unused_data
        members += 1
    assert members == 2 and out == plain
    assert np.array_equal(nifti.read(path, dtype=np.float32), a)
