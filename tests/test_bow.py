"""SURVEY.md 8(f) rank 2 -- BoW transform.  Golden vectors come from the reference's own pyDBoW classes
(tests/golden/make_golden.py::bow_case); the CPU test pins the oracle restatement, the GPU tests pin the kernel + the
Python assembly (including the reference's stale-node-id quirk and its float accumulation order)."""
import hashlib
import os
import types

import numpy as np
import pytest

from oracle.bow_py import Vocabulary, make_vocab_text


def _golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "bow_small.npz"))
    text = make_vocab_text()
    assert hashlib.sha256(text.encode()).hexdigest() == str(g["vocab_sha"]), "synthetic vocabulary drifted from the fixture"
    return g, Vocabulary.from_text(text)


def _check(bv, fv, g, lu):
    assert list(bv.keys()) == g[f"bv_keys_{lu}"].tolist()
    assert np.array_equal(np.array(list(bv.values()), np.float64), g[f"bv_vals_{lu}"])          # bit-exact doubles
    assert list(fv.keys()) == g[f"fv_keys_{lu}"].tolist()
    assert [len(v) for v in fv.values()] == g[f"fv_lens_{lu}"].tolist()
    assert [i for v in fv.values() for i in v] == g[f"fv_idx_{lu}"].tolist()


@pytest.mark.parametrize("lu", [4, 2, 1])
def test_oracle_restatement_vs_reference_pydbow(golden_dir, lu):
    g, voc = _golden(golden_dir)
    assert len(voc.children) == int(g["n_nodes"]) and voc.n_words == int(g["n_words"])
    bv, fv = voc.transform(g["desc"], lu)
    _check(bv, fv, g, lu)


def _as_reference_model(voc):
    """The data model of the reference's TemplatedVocabulary (.L, .nodes[i].children/.descriptor/.weight/.word_id)."""
    nodes = [types.SimpleNamespace(children=list(voc.children[i]), descriptor=None if i == 0 else voc.desc[i],
                                   weight=voc.weight[i], word_id=voc.word_id[i]) for i in range(len(voc.children))]
    return types.SimpleNamespace(L=voc.L, k=voc.k, nodes=nodes)


@pytest.mark.gpu
@pytest.mark.parametrize("lu", [4, 2, 1])
def test_gpu_transform_vs_reference_pydbow(golden_dir, lu):
    from pyorbslam_b200.bow import GpuVocabulary
    g, voc = _golden(golden_dir)
    gv = GpuVocabulary(_as_reference_model(voc))
    bv, fv = gv.transform(g["desc"], lu)
    _check(bv, fv, g, lu)
    bv, fv = gv.transform(np.zeros((0, 32), np.uint8), lu)
    assert bv == {} and fv == {}


@pytest.mark.gpu
def test_gpu_transform_resident_descriptors_and_install_hook():
    from pyorbslam_b200 import ORBextractor
    from pyorbslam_b200.bow import install_vocabulary
    from pyorbslam_b200.synthetic import make_stereo_pair
    voc = Vocabulary.from_text(make_vocab_text(seed=11, k=10, L=4, p_early_leaf=0.05))
    model = _as_reference_model(voc)
    model.transform = lambda f, lu=4: (_ for _ in ()).throw(AssertionError("unpatched"))
    gv = install_vocabulary(model)
    e = ORBextractor(2000, 1.2, 8, 20, 7)
    _, desc = e.operator_kd(make_stereo_pair(3)[0])
    bv, fv = model.transform(desc, 4)                # Frame.compute_BoW: self.mpORBvocabulary.transform(self.mDescriptors, 4)
    obv, ofv = voc.transform(desc, 4)
    assert list(bv.items()) == list(obv.items()) and list(fv.items()) == list(ofv.items())
    # the resident path (descriptor array identity) and the upload path agree
    leaf_r, lvl_r = gv.descend(desc, 4, extractor=e)
    leaf_h, lvl_h = gv.descend(desc.copy(), 4)
    assert np.array_equal(leaf_r, leaf_h) and np.array_equal(lvl_r, lvl_h)
    assert abs(sum(bv.values()) - 1.0) < 1e-12 and sum(len(v) for v in fv.values()) <= len(desc)
