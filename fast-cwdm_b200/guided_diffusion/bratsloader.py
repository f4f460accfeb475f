"""BraTS case loader with the reference's item layout (guided_diffusion/bratsloader.py:9-109): every leaf directory
under ``directory`` is one case whose files are named ``<anything>-<seq>.nii.gz`` with ``seq`` in t1n / t1c / t2w / t2f /
seg; an item is ``{'t1n','t1c','t2w','t2f': (1,224,224,160) float32 in [0,1] (or zeros(1) when the file is absent),
'missing': name or 'none', 'subj': path (eval / auto modes) or 'dummy_string', 'filedict': {...}}``.

Two differences, both about where the work happens:

* files are read with ``fcwdm.nifti`` (nibabel is not in this image);
* ``raw=True`` skips the host-side ``clip_and_normalize`` (two ``np.quantile`` sorts of 8.9 M voxels per modality,
  bratsloader.py:104-108) and returns the raw (240,240,155) float32 volumes: ``fcwdm.preprocess`` /
  ``VolumeStream(raw=True)`` then do clip + normalise + pad + crop on the GPU, so a case goes
  disk -> GPU without a host sort.  The default (``raw=False``) keeps the reference's host arithmetic for callers that
  expect normalised tensors from the loader (scripts/train.py).
"""
import os

import numpy as np
import torch
import torch.utils.data

from fcwdm import nifti

SEQTYPES = ('t1n', 't1c', 't2w', 't2f', 'seg')
MODALITIES = SEQTYPES[:4]


def _seqtype_of(filename):
    stem = filename
    for ext in (".gz", ".nii"):
        if stem.endswith(ext):
            stem = stem[:-len(ext)]
    tail = stem.rsplit('-', 1)[-1]
    return tail if tail in SEQTYPES else None


class BRATSVolumes(torch.utils.data.Dataset):
    def __init__(self, directory, mode='train', gen_type=None, raw=False):
        super().__init__()
        self.mode = mode
        self.directory = os.path.expanduser(directory)
        self.gentype = gen_type
        self.raw = raw
        self.seqtypes = list(SEQTYPES)
        self.seqtypes_set = set(SEQTYPES)
        self.database = []
        for root, dirs, files in sorted(os.walk(self.directory)):
            if dirs:                                        # only leaf directories hold cases
                continue
            case = {}
            for f in sorted(files):
                seq = _seqtype_of(f)
                if seq is not None:
                    case[seq] = os.path.join(root, f)
            if case:
                self.database.append(case)

    def __len__(self):
        return len(self.database)

    def __getitem__(self, x):
        filedict = self.database[x]
        item = {}
        missing = 'none'
        for seq in MODALITIES:
            if seq not in filedict:
                missing = seq
                item[seq] = torch.zeros(1)
                continue
            vol = nifti.read(filedict[seq], dtype=np.float32 if self.raw else np.float64)
            if self.raw:
                item[seq] = torch.from_numpy(np.ascontiguousarray(vol))
            else:
                item[seq] = pad_and_crop(torch.from_numpy(clip_and_normalize(vol)).float())
        if self.mode in ('eval', 'auto'):
            subj = filedict['t1n'] if 't1n' in filedict else filedict['t2f']
        else:
            subj = 'dummy_string'
        item.update(missing=missing, subj=subj, filedict=filedict)
        return item


def pad_and_crop(vol):
    """(240,240,155) -> (1,224,224,160): zero-pad the slice axis to 160, drop 8 voxels on each in-plane border."""
    out = torch.zeros(1, vol.shape[0], vol.shape[1], max(160, vol.shape[2]))
    out[0, :, :, :vol.shape[2]] = vol
    return out[:, 8:-8, 8:-8, :]


def clip_and_normalize(img):
    """Clip to the [0.1 %, 99.9 %] quantiles, then min-max to [0, 1] (host / numpy, as the reference's loader does)."""
    lo, hi = np.quantile(img, 0.001), np.quantile(img, 0.999)
    clipped = np.clip(img, lo, hi)
    cmin = clipped.min()
    return (clipped - cmin) / (clipped.max() - cmin)
