"""TEST INFRASTRUCTURE: the three nibabel names scripts/sample.py uses (`Nifti1Image`, `save`, `load`), backed by
fcwdm.nifti (nibabel is not installed in this image).  Installed as sys.modules['nibabel'] by the script-replay test."""
import types

import numpy as np

from fcwdm import nifti


class Nifti1Image:
    def __init__(self, dataobj, affine, header=None):
        self.dataobj = np.asarray(dataobj)
        self.affine = np.asarray(affine, dtype=np.float64)
        self.header = header

    def get_fdata(self):
        return np.asarray(self.dataobj, dtype=np.float64)


def save(img, filename):
    nifti.write(str(filename), np.asarray(img.dataobj), affine=img.affine)


def load(filename):
    data, hdr = nifti.read(str(filename), return_header=True)
    return Nifti1Image(data, hdr.affine, hdr)


def as_module():
    m = types.ModuleType("nibabel")
    m.Nifti1Image, m.save, m.load = Nifti1Image, save, load
    return m
