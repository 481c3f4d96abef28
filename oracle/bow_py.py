"""TEST INFRASTRUCTURE -- Python restatement of the reference's BoW transform:
pyDBoW/TemplatedVocabulary.py:44-82 (load_from_text_file), :108-131 (transform), :133-163 (transform_feature),
pyDBoW/FORB.py:31-33 (distance), pyDBoW/BowVector.py, pyDBoW/FeatureVector.py.
Pinned against vectors produced by the reference's own classes (tests/golden/make_golden.py -> bow_small.npz).

Reference quirks kept on purpose:
  * the node id handed to FeatureVector is only updated when the descent passes depth L - levelsup; a feature whose
    path is shorter inherits the node id of the PREVIOUS feature (transform passes node_id back in, :121-122, :157-158);
  * weights are accumulated with += in feature order, normalised by a sum taken in ascending word order."""
from collections import OrderedDict

import numpy as np


def make_vocab_text(seed=5, k=6, L=4, p_early_leaf=0.12, p_zero_weight=0.08):
    """A small synthetic vocabulary in the ORBvoc.txt text format: header `k L scoring weighting`, then one line per node
    `parent is_leaf d0 .. d31 weight`.  Some branches stop early (leaves above depth L) and some leaves weigh 0."""
    rng = np.random.default_rng(seed)
    lines = [f"{k} {L} 0 0"]
    frontier = [(0, 0)]          # (node id, depth)
    next_id = 1
    while frontier:
        nxt = []
        for parent, depth in frontier:
            for _ in range(k):
                nid = next_id
                next_id += 1
                leaf = depth + 1 == L or (depth + 1 >= 2 and rng.random() < p_early_leaf)
                desc = " ".join(str(int(v)) for v in rng.integers(0, 256, 32))
                w = 0.0 if (leaf and rng.random() < p_zero_weight) else float(np.round(rng.uniform(0.1, 9.0), 6))
                lines.append(f"{parent} {1 if leaf else 0} {desc} {w if leaf else 0.0}")
                if not leaf:
                    nxt.append((nid, depth + 1))
        frontier = nxt
    return "\n".join(lines) + "\n"


class Vocabulary:
    def __init__(self):
        self.k = self.L = 0
        self.children = [[]]
        self.desc = [np.zeros(32, np.int64)]
        self.weight = [0.0]
        self.word_id = [0]
        self.n_words = 0

    @classmethod
    def from_text(cls, text):
        v = cls()
        rows = text.strip().split("\n")
        head = rows[0].split()
        v.k, v.L = int(head[0]), int(head[1])
        for row in rows[1:]:
            parts = row.split()
            parent, is_leaf = int(parts[0]), int(parts[1])
            nid = len(v.children)
            v.children.append([])
            v.desc.append(np.array(list(map(float, parts[2:-1]))).astype(int))
            v.weight.append(float(parts[-1]))
            v.children[parent].append(nid)
            if is_leaf > 0:
                v.word_id.append(v.n_words)
                v.n_words += 1
            else:
                v.word_id.append(0)
        return v

    @staticmethod
    def distance(a, b):
        return sum(bin(int(x)).count("1") for x in np.bitwise_xor(a, b))

    def descend(self, feature, nid, levels_up):
        node, level = 0, 0
        target = self.L - levels_up
        while self.children[node]:
            kids = self.children[node]
            node = kids[0]
            best = self.distance(feature, self.desc[node])
            for c in kids[1:]:
                d = self.distance(feature, self.desc[c])
                if d < best:
                    best, node = d, c
            level += 1
            if nid is not None and level == target:
                nid = node
        return self.word_id[node], nid, self.weight[node], node

    def transform(self, features, levels_up=4):
        words, feats = {}, {}
        nid = 0
        for i in range(features.shape[0]):
            word, nid, w, _ = self.descend(features[i], nid, levels_up)
            if w > 0:
                words[word] = words[word] + w if word in words else w
                feats.setdefault(nid, []).append(i)
        words = OrderedDict(sorted(words.items())) if words else {}
        feats = OrderedDict(sorted(feats.items())) if feats else {}
        total = sum(words.values())
        if total > 0:
            for wid in words:
                words[wid] /= total
        return words, feats
