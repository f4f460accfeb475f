"""Minimal NIfTI-1 single-file reader / writer (``.nii`` and ``.nii.gz``) for the sampling driver and the BraTS loader.

The reference reads its inputs with ``nibabel.load(path).get_fdata()`` (bratsloader.py:46) and writes results with
``nib.save(nib.Nifti1Image(array, np.eye(4)), 'sample.nii.gz')`` (sample.py:141-145); nibabel is not available in
this image, and at >= 20 volumes/s per GPU its single-threaded gzip would be the bottleneck anyway.
This module implements the part of the published NIfTI-1 format those two calls use:

* header: the 348-byte ``nifti_1_header`` (either byte order on read), ``dim``, ``datatype``/``bitpix``, ``pixdim``,
  ``vox_offset``, ``scl_slope``/``scl_inter``, ``qform``/``sform`` codes, the sform rows, magic ``n+1``;
* data: voxel array in file order (first index fastest), the integer and float datatypes, optional linear scaling;
* ``read`` returns what ``get_fdata()`` would -- a float64 array (or ``dtype=`` of the caller's choice) with scaling
  applied, shape ``dim[1..ndim]`` -- plus the parsed header; ``write`` stores float32 (or the array's own supported dtype)
  with an affine in the sform (code 2, "aligned", as nibabel does for an image built from an array and an affine), a
  matching qform quaternion for a pure-diagonal/translation affine, and ``pixdim`` = the affine's column norms;
* gzip level 1 by default (nibabel's default) -- ``zlib`` releases the GIL, so several writer threads compress in
  parallel.
"""
import gzip
import os
import struct
import zlib

import numpy as np

HEADER_SIZE = 348
_DTYPES = {2: np.uint8, 4: np.int16, 8: np.int32, 16: np.float32, 64: np.float64, 256: np.int8, 512: np.uint16,
           768: np.uint32, 1024: np.int64, 1280: np.uint64}
_CODES = {np.dtype(v): k for k, v in _DTYPES.items()}


class NiftiError(ValueError):
    pass


class Header:
    """The fields of a nifti_1_header this package reads or writes."""

    def __init__(self):
        self.dim = (0,) * 8
        self.datatype = 16
        self.bitpix = 32
        self.pixdim = (1.0,) * 8
        self.vox_offset = 352.0
        self.scl_slope = 0.0
        self.scl_inter = 0.0
        self.qform_code = 0
        self.sform_code = 0
        self.quatern = (0.0, 0.0, 0.0)
        self.qoffset = (0.0, 0.0, 0.0)
        self.srow = np.eye(4)[:3].copy()
        self.xyzt_units = 0
        self.descrip = b""
        self.byteorder = "<"
        self.raw = None                     # the 348 bytes as read (None for a header built here)

    @property
    def shape(self):
        return tuple(int(v) for v in self.dim[1:1 + int(self.dim[0])])

    @property
    def affine(self):
        """sform if set, else the qform's scaling/translation part, else diag(pixdim)."""
        a = np.eye(4)
        if self.sform_code > 0:
            a[:3] = self.srow
            return a
        if self.qform_code > 0:
            b, c, d = self.quatern
            aa = max(0.0, 1.0 - (b * b + c * c + d * d)) ** 0.5
            R = np.array([[aa * aa + b * b - c * c - d * d, 2 * (b * c - aa * d), 2 * (b * d + aa * c)],
                          [2 * (b * c + aa * d), aa * aa + c * c - b * b - d * d, 2 * (c * d - aa * b)],
                          [2 * (b * d - aa * c), 2 * (c * d + aa * b), aa * aa + d * d - b * b - c * c]])
            qfac = -1.0 if self.pixdim[0] < 0 else 1.0
            a[:3, :3] = R * np.array([self.pixdim[1], self.pixdim[2], self.pixdim[3] * qfac])
            a[:3, 3] = self.qoffset
            return a
        a[0, 0], a[1, 1], a[2, 2] = self.pixdim[1:4]
        return a


def _open_bytes(path):
    with open(path, "rb") as f:
        data = f.read()
    if data[:2] == b"\x1f\x8b":
        data = gzip.decompress(data)
    return data


def parse_header(buf):
    if len(buf) < HEADER_SIZE:
        raise NiftiError("file shorter than a NIfTI-1 header")
    for order in ("<", ">"):
        if struct.unpack_from(order + "i", buf, 0)[0] == HEADER_SIZE:
            break
    else:
        raise NiftiError("sizeof_hdr is not 348: not a NIfTI-1 file")
    magic = bytes(buf[344:348])
    if magic not in (b"n+1\x00", b"ni1\x00"):
        raise NiftiError(f"bad NIfTI-1 magic {magic!r}")
    if magic == b"ni1\x00":
        raise NiftiError("two-file (.hdr/.img) NIfTI pairs are not supported")
    h = Header()
    h.byteorder = order
    h.raw = bytes(buf[:HEADER_SIZE])
    h.dim = struct.unpack_from(order + "8h", buf, 40)
    h.datatype, h.bitpix = struct.unpack_from(order + "2h", buf, 70)
    h.pixdim = struct.unpack_from(order + "8f", buf, 76)
    h.vox_offset, h.scl_slope, h.scl_inter = struct.unpack_from(order + "3f", buf, 108)
    h.xyzt_units = buf[123]
    h.descrip = bytes(buf[148:228]).rstrip(b"\x00")
    h.qform_code, h.sform_code = struct.unpack_from(order + "2h", buf, 252)
    h.quatern = struct.unpack_from(order + "3f", buf, 256)
    h.qoffset = struct.unpack_from(order + "3f", buf, 268)
    h.srow = np.array(struct.unpack_from(order + "12f", buf, 280), dtype=np.float64).reshape(3, 4)
    if not 1 <= h.dim[0] <= 7:
        raise NiftiError(f"dim[0] = {h.dim[0]} out of range")
    if h.datatype not in _DTYPES:
        raise NiftiError(f"unsupported NIfTI datatype code {h.datatype}")
    return h


def read(path, dtype=np.float64, return_header=False):
    """The image array as ``nibabel.load(path).get_fdata(dtype=dtype)`` returns it: shape dim[1..n], scaling applied
    (``scl_slope`` 0 or NaN means none), Fortran-ordered like nibabel's (a transposed view of the file-order buffer)."""
    buf = _open_bytes(path)
    h = parse_header(buf)
    shape = h.shape
    n = int(np.prod(shape))
    dt = np.dtype(_DTYPES[h.datatype]).newbyteorder(h.byteorder)
    off = int(h.vox_offset)
    if off < 352:
        raise NiftiError(f"vox_offset {h.vox_offset} inside the header")
    if len(buf) < off + n * dt.itemsize:
        raise NiftiError("file is shorter than its header claims")
    data = np.frombuffer(buf, dtype=dt, count=n, offset=off).reshape(shape[::-1]).T     # first index fastest
    slope, inter = float(h.scl_slope), float(h.scl_inter)
    if slope != 0.0 and np.isfinite(slope) and (slope != 1.0 or inter != 0.0):
        out = data.astype(dtype) * slope + (inter if np.isfinite(inter) else 0.0)
    else:
        out = data.astype(dtype)
    return (out, h) if return_header else out


def _is_positive_diagonal(affine):
    """True when the 3x3 part is a positive diagonal (identity rotation): the qform quaternion is then (0, 0, 0) and
    the offsets are the translation; with rotation or flips only the sform is written (qform_code stays 0)."""
    M = affine[:3, :3]
    return not np.count_nonzero(M - np.diag(np.diag(M))) and bool(np.all(np.diag(M) > 0))


def build_header(shape, dtype, affine=None, like=None):
    """348 header bytes + 4 extension bytes for an array of `shape` / `dtype`.  `like` (a Header read from another file)
    donates orientation, units and description -- the 'header copied from a conditioning modality' case."""
    dtype = np.dtype(dtype)
    if dtype not in _CODES:
        raise NiftiError(f"dtype {dtype} has no NIfTI-1 code")
    if not 1 <= len(shape) <= 7:
        raise NiftiError("NIfTI-1 stores 1 to 7 dimensions")
    if max(shape) > 32767:
        raise NiftiError("NIfTI-1 dims are int16")
    from_like = affine is None and like is not None
    affine = like.affine if from_like else (np.eye(4) if affine is None else np.asarray(affine, dtype=np.float64))
    b = bytearray(HEADER_SIZE + 4)
    struct.pack_into("<i", b, 0, HEADER_SIZE)
    struct.pack_into("<8h", b, 40, len(shape), *(list(shape) + [1] * (7 - len(shape))))
    struct.pack_into("<2h", b, 70, _CODES[dtype], dtype.itemsize * 8)
    norms = np.sqrt((affine[:3, :3] ** 2).sum(axis=0))
    pixdim = [1.0] + [float(v) for v in norms] + [1.0] * 4
    plain = _is_positive_diagonal(affine)
    qform_code = 0
    if from_like:
        pixdim = list(like.pixdim)
    struct.pack_into("<3f", b, 108, float(HEADER_SIZE + 4), 1.0, 0.0)          # vox_offset, scl_slope, scl_inter
    if like is not None:
        b[123] = like.xyzt_units
        b[148:148 + len(like.descrip[:79])] = like.descrip[:79]
        if like.qform_code > 0 and from_like:
            qform_code = like.qform_code
            struct.pack_into("<3f", b, 256, *like.quatern)
            struct.pack_into("<3f", b, 268, *like.qoffset)
    if plain and qform_code == 0:
        struct.pack_into("<3f", b, 256, 0.0, 0.0, 0.0)
        struct.pack_into("<3f", b, 268, *(float(v) for v in affine[:3, 3]))
    struct.pack_into("<8f", b, 76, *pixdim)
    struct.pack_into("<2h", b, 252, qform_code, 2)                           # sform_code 2 = aligned
    struct.pack_into("<12f", b, 280, *(float(v) for v in affine[:3].reshape(-1)))
    b[344:348] = b"n+1\x00"
    return bytes(b)


def encode(array, affine=None, like=None, compresslevel=1, gz=True):
    """The bytes of a .nii(.gz) file holding `array` (float32 unless the array already has a supported dtype)."""
    a = np.asarray(array)
    if a.dtype not in _CODES or a.dtype == np.float64:
        a = a.astype(np.float32)
    payload = build_header(a.shape, a.dtype, affine, like) + np.asfortranarray(a).tobytes(order="F")
    if not gz:
        return payload
    return gzip_members(payload, compresslevel)


_GZ_CHUNK = 4 << 20
_gz_pool = None


def gzip_members(payload, compresslevel=1, threads=None):
    """gzip `payload` as a sequence of independent members of <= 4 MB, deflated concurrently (zlib releases the GIL).
    The concatenation of gzip members is a valid gzip file (RFC 1952 section 2.2) that gzip / zlib / nibabel read as one
    stream; a 35 MB volume costs one core ~0.35 s at level 1 as a single member and ~1/8 of that on 8 threads -- the
    writer side of fcwdm.sample_driver was bound by exactly this."""
    global _gz_pool
    view = memoryview(payload)
    if len(view) <= _GZ_CHUNK:
        co = zlib.compressobj(compresslevel, zlib.DEFLATED, 16 + zlib.MAX_WBITS)
        return co.compress(view) + co.flush()

    def one(lo):
        co = zlib.compressobj(compresslevel, zlib.DEFLATED, 16 + zlib.MAX_WBITS)
        return co.compress(view[lo:lo + _GZ_CHUNK]) + co.flush()

    if _gz_pool is None:
        from concurrent.futures import ThreadPoolExecutor
        _gz_pool = ThreadPoolExecutor(max_workers=threads or min(8, os.cpu_count() or 1), thread_name_prefix="fcwdm-gz")
    return b"".join(_gz_pool.map(one, range(0, len(view), _GZ_CHUNK)))


def write(path, array, affine=None, like=None, compresslevel=1):
    data = encode(array, affine, like, compresslevel, gz=str(path).endswith(".gz"))
    tmp = f"{path}.part"
    with open(tmp, "wb") as f:
        f.write(data)
    os.replace(tmp, path)
    return len(data)
