"""GPU BoW transform -- SURVEY.md 8(f) rank 2: `TemplatedVocabulary.transform(descriptors, levelsup)`
(reference pyDBoW/TemplatedVocabulary.py:108-163), the direct consumer of the extractor's descriptors through
`Frame.compute_BoW` (Frame.py:123-125).  The reference walks the tree in pure Python (one `bin().count` Hamming distance
per child per level per feature); here the descent is one kernel (`k_vocab_descend`, warp per feature) and only the
dictionary assembly stays in Python, reproducing the reference's accumulation order and its node-id quirk exactly.

    import pyorbslam_b200.bow as bow
    bow.install_vocabulary(voc)          # voc = the reference's TemplatedVocabulary after load_from_text_file(...)
    # Frame.compute_BoW() -> voc.transform(self.mDescriptors, 4) now runs on the GPU, same return value
"""
import ctypes as C
from collections import OrderedDict

import numpy as np

from . import _lib
from .extractor import ORBextractor


class GpuVocabulary:
    """Wraps any object with the reference vocabulary's data model: `.L` and `.nodes[i]` having `.children` (ids in order),
    `.descriptor` (32 values, None for the root), `.weight`, `.word_id`."""

    def __init__(self, voc, device=0):
        self._h = None
        nodes = voc.nodes
        n = len(nodes)
        self.L = int(voc.L)
        begin = np.zeros(n + 1, np.int32)
        ids = []
        desc = np.zeros((n, 32), np.uint8)
        self.weight = [0.0] * n
        self.word_id = [0] * n
        for i, nd in enumerate(nodes):
            ids.extend(int(c) for c in nd.children)
            begin[i + 1] = len(ids)
            if nd.descriptor is not None:
                desc[i] = np.asarray(nd.descriptor).astype(np.int64) & 0xff
            self.weight[i] = nd.weight
            self.word_id[i] = nd.word_id
        ids = np.asarray(ids, np.int32)
        l = _lib.lib()
        l.b200orb_vocab_create.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]
        l.b200orb_vocab_destroy.argtypes = [C.c_void_p]
        l.b200orb_vocab_destroy.restype = None
        l.b200orb_vocab_transform.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        l.b200orb_vocab_transform_resident.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        h = C.c_void_p()
        _lib.check(l.b200orb_vocab_create(n, begin.ctypes.data, ids.ctypes.data if len(ids) else None, desc.ctypes.data, int(device), C.byref(h)))
        self._h = h

    def __del__(self):
        if getattr(self, "_h", None):
            try:
                _lib.lib().b200orb_vocab_destroy(self._h)
            except Exception:
                pass
            self._h = None

    def descend(self, features, levels_up=4, extractor=None):
        """-> (leaf node id, node id at depth L - levels_up or -1) per feature.  With `extractor`, the descriptors of its
        last operator_kd are read where they already are (device memory)."""
        nid_level = self.L - int(levels_up)
        if extractor is not None:
            n = extractor._last_n
            leaf, lvl = np.empty(max(n, 0), np.int32), np.empty(max(n, 0), np.int32)
            if n > 0:
                _lib.check(_lib.lib().b200orb_vocab_transform_resident(self._h, extractor._h, nid_level, leaf.ctypes.data, lvl.ctypes.data))
            return leaf, lvl
        f = np.ascontiguousarray(features, np.uint8).reshape(-1, 32) if len(features) else np.zeros((0, 32), np.uint8)
        leaf, lvl = np.empty(len(f), np.int32), np.empty(len(f), np.int32)
        if len(f):
            _lib.check(_lib.lib().b200orb_vocab_transform(self._h, f.ctypes.data, len(f), nid_level, leaf.ctypes.data, lvl.ctypes.data))
        return leaf, lvl

    def transform(self, features, levels_up=4):
        """Same return value as TemplatedVocabulary.transform: (word id -> L1-normalised weight, node id -> feature indices),
        both ordered by key."""
        ext = next((e for e in list(ORBextractor._live) if e._last_desc is features and e._last_n == len(features)), None)
        leaf, lvl = self.descend(features, levels_up, extractor=ext)
        words, feats = {}, {}
        nid = 0
        weight, word_id = self.weight, self.word_id
        for i, (lf, lv) in enumerate(zip(leaf.tolist(), lvl.tolist())):
            if lv >= 0:
                nid = lv                      # otherwise the previous feature's node id is kept (reference behaviour, :121-122)
            w = weight[lf]
            if w > 0:
                wid = word_id[lf]
                words[wid] = words[wid] + w if wid in words else w
                if nid in feats:
                    feats[nid].append(i)
                else:
                    feats[nid] = [i]
        words = OrderedDict(sorted(words.items())) if words else {}
        feats = OrderedDict(sorted(feats.items())) if feats else {}
        total = sum(words.values())
        if total > 0:
            for wid in words:
                words[wid] /= total
        return words, feats


def install_vocabulary(voc, device=0):
    """voc.transform = GPU transform (instance attribute; the class and every other method stay the reference's)."""
    g = GpuVocabulary(voc, device)
    voc._b200orb_gpu = g
    voc.transform = g.transform
    return g
