"""Parameter-pack cache shared by the module mirrors.

Kernels read packed device blobs, not the nn.Parameters; a blob is rebuilt when any
source parameter changed (``_version`` bumps on optimizer steps / ``load_state_dict``)
or was re-allocated (``data_ptr``)."""
import torch


class PackCache:
    def __init__(self):
        self._key = None
        self._val = None

    @staticmethod
    def key_of(tensors):
        return tuple((t.data_ptr(), t._version, t.device) for t in tensors)

    def get(self, tensors, builder):
        k = self.key_of(tensors)
        if k != self._key:
            with torch.no_grad():
                self._val = builder()
            self._key = k
        return self._val

    def invalidate(self):
        self._key = None
        self._val = None
