"""Index sets: drop-in for the reference's src/IndexSetGenerator.py (host side, bit-exact contract).

The reference materialises every region as a Python list of ints (IndexSetGenerator.py:51-92); all of them
are contiguous ranges, so the native representation here is an int64 offsets array per layer
(`.offsets[j]`, length R_j + 1).  `.index_set[j][l]` is still available (built lazily, as `range`-backed
lists) for code that reads it."""
import numpy as np


class IndexSetUniform(object):
    def __init__(self, sample_length, resolution, divider, n_regions=None,
                 min_percentage_of_samples_per_region=None):
        # IndexSetGenerator.py:5-49
        self.resolution = int(resolution)
        if n_regions is None:
            self.divider = 0 if self.resolution == 0 else int(divider)
        self.sample_length = int(sample_length)
        self.offsets = []
        self._index_set = None
        if n_regions is None:
            for m in range(self.resolution + 1):
                self.offsets.append(self._get_offsets(m))
        else:
            self.region_ind = []
            self.min_number_of_samples_per_region = []
            for m in range(self.resolution + 1):
                n_regions_res = n_regions[m]
                percentage = 0.25 if min_percentage_of_samples_per_region is None \
                    else min_percentage_of_samples_per_region
                self.min_number_of_samples_per_region.append(
                    int(np.floor(np.divide(self.sample_length, n_regions_res) * percentage)))
                off, region_ind = self._get_offsets_random(m, n_regions_res)
                self.offsets.append(off)
                self.region_ind.append(region_ind)

    def get_n_resolutions(self):
        return self.resolution

    def get_index_set(self, resolution):
        return self.index_set[int(resolution)]

    @property
    def index_set(self):
        if self._index_set is None:
            self._index_set = [[list(range(int(off[l]), int(off[l + 1]))) for l in range(len(off) - 1)]
                               for off in self.offsets]
        return self._index_set

    def _get_offsets(self, resolution):
        # IndexSetGenerator.py:51-65
        number_of_regions = int(np.power(self.divider, resolution))
        samples_per_region = self.sample_length // number_of_regions
        if samples_per_region < 1:
            raise ValueError('*** Chosen resolution is too large! ***')
        off = np.arange(number_of_regions + 1, dtype=np.int64) * samples_per_region
        off[-1] = self.sample_length
        return off

    def _get_offsets_random(self, resolution, number_of_regions):
        # IndexSetGenerator.py:67-92 (consumes the global NumPy RNG exactly like the reference)
        sample_length = self.sample_length
        if number_of_regions == 1:
            return np.array([0, sample_length], dtype=np.int64), None
        repeat_flg = 1
        while repeat_flg != 0:
            region_ind = [0]
            for _ in range(number_of_regions - 1):
                region_ind.append(np.random.randint(1, sample_length))
            region_ind.append(sample_length)
            region_ind = np.sort(region_ind)
            diff_ = np.diff(region_ind)
            repeat_flg = int(np.sum(diff_ < self.min_number_of_samples_per_region[resolution]))
        return np.asarray(region_ind, dtype=np.int64), region_ind


def offsets_of(index_set_obj):
    """Region offsets per layer of any index-set object exposing the reference's `.index_set`."""
    if hasattr(index_set_obj, 'offsets'):
        return [np.asarray(o, dtype=np.int64) for o in index_set_obj.offsets]
    layers = []
    for regions in index_set_obj.index_set:
        off = [int(regions[0][0])]
        for r in regions:
            if len(r) == 0 or int(r[0]) != off[-1] or int(r[-1]) - int(r[0]) + 1 != len(r):
                raise ValueError('index sets must be contiguous, ordered ranges (as IndexSetGenerator builds them)')
            off.append(int(r[-1]) + 1)
        layers.append(np.asarray(off, dtype=np.int64))
    return layers
