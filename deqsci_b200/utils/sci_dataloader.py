"""CACTI benchmark loaders — same classes and tensor conventions as the reference's
utils/sci_dataloader.py:218-274 ({'gt': orig/255 [H,W,F], 'mask': [H,W,T], 'meas': meas/255 [H,W,M]},
float32), without the scipy private names the reference imports (:10-11, removed upstream)."""
import os

import numpy as np
import scipy.io as sio
from torch.utils.data import Dataset


def directory_filelist(target_directory):
    return sorted(f for f in os.listdir(target_directory)
                  if os.path.isfile(os.path.join(target_directory, f)) and not f.startswith('.'))


def _loadmat(path):
    """MATLAB v5-v7.2 through scipy; v7.3 (HDF5) through h5py when it is installed."""
    try:
        return sio.loadmat(path), False
    except NotImplementedError:
        import h5py  # noqa: only needed for -v7.3 files
        f = h5py.File(path, 'r')
        return {k: np.array(f[k]).transpose() for k in f.keys()}, True


def load_test_data(matfile):
    file, _ = _loadmat(matfile)
    return {'gt': np.float32(file['orig']) / 255, 'mask': np.float32(file['mask']),
            'meas': np.float32(file['meas']) / 255}


def load_mat(location, key):
    file, _ = _loadmat(location)
    if key == 'gt':
        for name in ('patch_save', 'p1', 'p2', 'p3'):
            if name in file:
                return np.float32(file[name] / 255)
        raise KeyError("no ground-truth variable in %s" % location)
    if key == 'meas':
        return np.float32(file['meas'] / 255)
    if key == 'mask':
        return np.float32(file['mask'])
    raise KeyError(key)


class SCITestDataset(Dataset):
    def __init__(self, dir):
        self.dir = dir
        self.filelist = directory_filelist(dir)

    def __len__(self):
        return len(self.filelist)

    def __getitem__(self, item):
        data = load_test_data(os.path.join(self.dir, self.filelist[item]))
        data['file'] = self.filelist[item]
        return data


class SCITrainingDatasetSubset(Dataset):
    def __init__(self, gt_directory, meas_directory, mask_location):
        names = directory_filelist(gt_directory)
        self.full_gt_filelist = [os.path.join(gt_directory, n) for n in names]
        self.full_meas_filelist = [os.path.join(meas_directory, n) for n in names]
        self.mask = load_mat(mask_location, 'mask')

    def __len__(self):
        return len(self.full_gt_filelist)

    def __getitem__(self, item):
        return {'gt': load_mat(self.full_gt_filelist[item], 'gt'), 'mask': self.mask,
                'meas': load_mat(self.full_meas_filelist[item], 'meas')}
