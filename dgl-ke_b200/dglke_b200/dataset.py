"""On-disk knowledge-graph readers (reference: dataloader/KGDataset.py:73-145 base reader, :186-397 built-in layouts,
:505-736 user-defined formats, :738-771 get_dataset).

Same file conventions, attributes and error conditions as the reference's classes, so that `dglke_train`'s data flags
(--data_path --dataset --format --data_files --delimiter --has_edge_importance) mean the same thing here.  Differences:

  * the files are parsed column-wise (pandas C parser / numpy), not line by line in Python -- Freebase's train.txt is
    304 M lines;
  * there is no network in this environment: a built-in dataset must already be unpacked under <data_path>/<name>/
    (the layout the reference's downloader produces); otherwise FileNotFoundError says where it was expected, and
    `dglke_b200.train` falls back to a synthetic graph of the dataset's published shape;
  * OGB-packaged datasets (wikikg2, biokg, wikikg90M: KGDataset.py:399-503) need the `ogb` package and are not read here.

get_dataset(...) returns an object with n_entities, n_relations, entity2id, relation2id (dict or None), train / valid /
test = (heads, rels, tails[, importance]) int64 arrays, emap_fname, rmap_fname -- what train.py:62-116 consumes."""
import os

import numpy as np

_FORMATS = {"hrt": (0, 1, 2), "htr": (0, 2, 1), "rht": (1, 0, 2), "rth": (2, 0, 1), "thr": (1, 2, 0), "trh": (2, 1, 0)}


def parse_srd_format(fmt):
    """'hrt' ... -> column of (head, relation, tail) in a triple line (KGDataset.py:53-65)."""
    if fmt not in _FORMATS:
        raise ValueError("unknown triple format %r (one of %s)" % (fmt, sorted(_FORMATS)))
    return _FORMATS[fmt]


def _read_columns(path, delimiter, skip_first_line=False, ncols=None):
    """All columns of a delimited text file as a list of 1-D object/str arrays."""
    import pandas as pd
    df = pd.read_csv(path, sep=delimiter, header=None, skiprows=1 if skip_first_line else 0, dtype=str,
                     keep_default_na=False, na_filter=False, engine="c", quoting=3, skip_blank_lines=True)
    if ncols is not None and df.shape[1] < ncols:
        raise ValueError("%s: expected at least %d %r-separated columns, found %d" % (path, ncols, delimiter, df.shape[1]))
    return [df[c].str.strip().to_numpy() for c in df.columns]


def _map_names(names, mapping, what, path):
    """vectorised dict lookup; KeyError (as the reference's dict indexing raises) names the first unknown key"""
    import pandas as pd
    ids = pd.Series(names).map(mapping)
    if ids.isna().any():
        raise KeyError("%s %r of %s is not in the %s dictionary" % (what, names[int(np.flatnonzero(ids.isna().to_numpy())[0])], path, what))
    return ids.to_numpy(dtype=np.int64)


class KGDataset:
    """entities / relations dictionaries ('id<delim>name' per line) + triple files by NAME (KGDataset.py:73-145)."""

    def __init__(self, entity_path, relation_path, train_path, valid_path=None, test_path=None, format=(0, 1, 2),
                 delimiter="\t", skip_first_line=False):
        self.delimiter = delimiter
        self.entity2id, self.n_entities = self.read_entity(entity_path)
        self.relation2id, self.n_relations = self.read_relation(relation_path)
        self.train = self.read_triple(train_path, "train", skip_first_line, format)
        self.valid = self.read_triple(valid_path, "valid", skip_first_line, format) if valid_path is not None else None
        self.test = self.read_triple(test_path, "test", skip_first_line, format) if test_path is not None else None

    def _read_dict(self, path):
        ids, names = _read_columns(path, self.delimiter, ncols=2)[:2]
        return dict(zip(names.tolist(), ids.astype(np.int64).tolist()))

    def read_entity(self, entity_path):
        m = self._read_dict(entity_path)
        return m, len(m)

    def read_relation(self, relation_path):
        m = self._read_dict(relation_path)
        return m, len(m)

    def read_triple(self, path, mode, skip_first_line=False, format=(0, 1, 2)):
        if path is None:
            return None
        print("Reading {} triples....".format(mode))
        cols = _read_columns(path, self.delimiter, skip_first_line, ncols=3)
        heads = _map_names(cols[format[0]], self.entity2id, "entity", path)
        rels = _map_names(cols[format[1]], self.relation2id, "relation", path)
        tails = _map_names(cols[format[2]], self.entity2id, "entity", path)
        print("Finished. Read {} {} triples.".format(len(heads), mode))
        return heads, rels, tails


class _BuiltIn(KGDataset):
    """<path>/<name>/{entities.dict, relations.dict, train.txt, valid.txt, test.txt} (FB15k, FB15k-237, wn18, wn18rr:
    KGDataset.py:186-331).  No download here."""
    FILES = ("entities.dict", "relations.dict", "train.txt", "valid.txt", "test.txt")

    def __init__(self, path, name):
        self.name = name
        self.path = os.path.join(path, name)
        missing = [f for f in self.FILES if not os.path.exists(os.path.join(self.path, f))]
        if missing:
            raise FileNotFoundError("built-in dataset %s: %s not found under %s (no network here: unpack "
                                    "https://data.dgl.ai/dataset/%s.zip there)" % (name, ", ".join(missing), self.path, name))
        super().__init__(*(os.path.join(self.path, f) for f in self.FILES))

    @property
    def emap_fname(self):
        return self.FILES[0]

    @property
    def rmap_fname(self):
        return self.FILES[1]


class KGDatasetFB15k(_BuiltIn):
    def __init__(self, path, name="FB15k"):
        super().__init__(path, name)


class KGDatasetFB15k237(_BuiltIn):
    def __init__(self, path, name="FB15k-237"):
        super().__init__(path, name)


class KGDatasetWN18(_BuiltIn):
    def __init__(self, path, name="wn18"):
        super().__init__(path, name)


class KGDatasetWN18rr(_BuiltIn):
    def __init__(self, path, name="wn18rr"):
        super().__init__(path, name)


class KGDatasetFreebase(_BuiltIn):
    """Full Freebase (KGDataset.py:333-397): the dictionaries' first line is the COUNT, the triples are integer ids in
    the order head, tail, relation."""
    FILES = ("entity2id.txt", "relation2id.txt", "train.txt", "valid.txt", "test.txt")

    def __init__(self, path, name="Freebase"):
        super().__init__(path, name)

    @staticmethod
    def _count(path):
        with open(path) as f:
            return int(f.readline().strip())

    def read_entity(self, entity_path):
        return None, self._count(entity_path)

    def read_relation(self, relation_path):
        return None, self._count(relation_path)

    def read_triple(self, path, mode, skip_first_line=False, format=None):
        if path is None:
            return None
        print("Reading {} triples....".format(mode))
        import pandas as pd
        a = pd.read_csv(path, sep=self.delimiter, header=None, skiprows=1 if skip_first_line else 0, dtype=np.int64,
                        engine="c").to_numpy()
        print("Finished. Read {} {} triples.".format(len(a), mode))
        return np.ascontiguousarray(a[:, 0]), np.ascontiguousarray(a[:, 2]), np.ascontiguousarray(a[:, 1])


def _check_files(path, files):
    for f in files:
        assert os.path.exists(os.path.join(path, f)), "File {} not exist in {}".format(f, path)


def _warn_delimiter(delimiter):
    if delimiter not in ["\t", "|", ",", ";"]:
        print("WARNING: delimiter {} is not in '\\t', '|', ',', ';'This is not tested by the developer".format(delimiter))


def _importance(cols, path):
    if len(cols) < 4:
        raise ValueError("%s: --has_edge_importance needs a 4th column" % path)
    w = cols[3].astype(np.float64)
    assert np.min(w) > 0.0, "Edge importance score should > 0"
    return w


class KGDatasetUDDRaw(KGDataset):
    """raw_udd_{hrt..}: triples by NAME in 1 (train) or 3 (train, valid, test) files; the dictionaries are built from the
    files in order of first appearance -- source before destination before relation, line by line -- and written to
    <path>/entities.tsv and <path>/relations.tsv as 'id<delim>name' (KGDataset.py:505-624)."""

    def __init__(self, path, name, delimiter, files, format, has_edge_importance=False):
        self.name = name
        _check_files(path, files)
        assert len(format) == 3
        fmt = parse_srd_format(format)
        self.delimiter = delimiter
        self.load_entity_relation(path, delimiter, files, fmt)
        assert len(files) == 1 or len(files) == 3, "raw_udd_{htr} format requires 1 or 3 input files. " \
            "When 1 files are provided, they must be train_file. " \
            "When 3 files are provided, they must be train_file, valid_file and test_file."
        _warn_delimiter(delimiter)
        self.has_edge_importance = has_edge_importance
        paths = [os.path.join(path, f) for f in files] + [None] * (3 - len(files))
        super().__init__("entities.tsv", "relation.tsv", paths[0], paths[1], paths[2], format=fmt, delimiter=delimiter)

    def load_entity_relation(self, path, delimiter, files, format):
        import pandas as pd
        ent_keys, rel_keys = [], []
        for fi in files:
            cols = _read_columns(os.path.join(path, fi), delimiter, ncols=3)
            src, rel, dst = cols[format[0]], cols[format[1]], cols[format[2]]
            # the reference assigns ids while walking the lines: src of line 0, dst of line 0, src of line 1, ...
            inter = np.empty(2 * len(src), dtype=object)
            inter[0::2], inter[1::2] = src, dst
            ent_keys.append(inter)
            rel_keys.append(rel)
        ents = pd.unique(np.concatenate(ent_keys)) if ent_keys else np.array([], dtype=object)
        rels = pd.unique(np.concatenate(rel_keys)) if rel_keys else np.array([], dtype=object)
        self.entity2id = {k: i for i, k in enumerate(ents.tolist())}
        self.relation2id = {k: i for i, k in enumerate(rels.tolist())}
        self.n_entities, self.n_relations = len(self.entity2id), len(self.relation2id)
        with open(os.path.join(path, "entities.tsv"), "w+") as f:
            f.writelines("{}{}{}\n".format(v, delimiter, k) for k, v in self.entity2id.items())
        with open(os.path.join(path, "relations.tsv"), "w+") as f:
            f.writelines("{}{}{}\n".format(v, delimiter, k) for k, v in self.relation2id.items())

    def read_entity(self, entity_path):
        return self.entity2id, self.n_entities

    def read_relation(self, relation_path):
        return self.relation2id, self.n_relations

    def read_triple(self, path, mode, skip_first_line=False, format=(0, 1, 2)):
        if path is None:
            return None
        print("Reading {} triples....".format(mode))
        cols = _read_columns(path, self.delimiter, skip_first_line, ncols=3)
        heads = _map_names(cols[format[0]], self.entity2id, "entity", path)
        rels = _map_names(cols[format[1]], self.relation2id, "relation", path)
        tails = _map_names(cols[format[2]], self.entity2id, "entity", path)
        print("Finished. Read {} {} triples.".format(len(heads), mode))
        if self.has_edge_importance:
            return heads, rels, tails, _importance(cols, path)
        return heads, rels, tails

    @property
    def emap_fname(self):
        return "entities.tsv"

    @property
    def rmap_fname(self):
        return "relations.tsv"


class KGDatasetUDD(KGDataset):
    """udd_{hrt..}: integer ids; files = entity map, relation map, train [, valid, test]; the maps are only COUNTED
    (one entity / relation per line) (KGDataset.py:626-736)."""

    def __init__(self, path, name, delimiter, files, format, has_edge_importance=False):
        self.name = name
        _check_files(path, files)
        fmt = parse_srd_format(format)
        assert len(files) == 3 or len(files) == 5, "udd_{htr} format requires 3 or 5 input files. " \
            "When 3 files are provided, they must be entity2id, relation2id, train_file. " \
            "When 5 files are provided, they must be entity2id, relation2id, train_file, valid_file and test_file."
        _warn_delimiter(delimiter)
        self.has_edge_importance = has_edge_importance
        p = [os.path.join(path, f) for f in files] + [None] * (5 - len(files))
        super().__init__(p[0], p[1], p[2], p[3], p[4], format=fmt, delimiter=delimiter)
        self.emap_file, self.rmap_file = files[0], files[1]

    @staticmethod
    def _lines(path):
        n = 0
        with open(path, "rb") as f:
            for _ in f:
                n += 1
        return n

    def read_entity(self, entity_path):
        return None, self._lines(entity_path)

    def read_relation(self, relation_path):
        return None, self._lines(relation_path)

    def read_triple(self, path, mode, skip_first_line=False, format=(0, 1, 2)):
        if path is None:
            return None
        print("Reading {} triples....".format(mode))
        cols = _read_columns(path, self.delimiter, skip_first_line, ncols=3)
        try:
            heads, rels, tails = (cols[format[k]].astype(np.int64) for k in range(3))
        except ValueError:
            print("For User Defined Dataset, both node ids and relation ids in the triplets should be int")
            raise
        print("Finished. Read {} {} triples.".format(len(heads), mode))
        assert np.max(heads) < self.n_entities, "Head node ID should not exceeds the number of entities {}".format(self.n_entities)
        assert np.max(tails) < self.n_entities, "Tail node ID should not exceeds the number of entities {}".format(self.n_entities)
        assert np.max(rels) < self.n_relations, "Relation ID should not exceeds the number of relations {}".format(self.n_relations)
        assert np.min(heads) >= 0, "Head node ID should >= 0"
        assert np.min(tails) >= 0, "Tail node ID should >= 0"
        assert np.min(rels) >= 0, "Relation ID should >= 0"
        if self.has_edge_importance:
            return heads, rels, tails, _importance(cols, path)
        return heads, rels, tails

    @property
    def emap_fname(self):
        return self.emap_file

    @property
    def rmap_fname(self):
        return self.rmap_file


_BUILT_IN = {"Freebase": KGDatasetFreebase, "FB15k": KGDatasetFB15k, "FB15k-237": KGDatasetFB15k237, "wn18": KGDatasetWN18,
             "wn18rr": KGDatasetWN18rr}


def get_dataset(data_path, data_name, format_str, delimiter="\t", files=None, has_edge_importance=False):
    """KGDataset.py:738-771."""
    if format_str == "built_in":
        if data_name in ("wikikg2", "biokg", "wikikg90M"):
            raise NotImplementedError("%s is packaged by OGB (KGDataset.py:399-503); the `ogb` package is not available here" % data_name)
        assert data_name in _BUILT_IN, "Unknown dataset {}".format(data_name)
        return _BUILT_IN[data_name](data_path)
    if format_str.startswith("raw_udd"):
        assert data_name != "FB15k", "You should provide the dataset name for raw_udd format."
        return KGDatasetUDDRaw(data_path, data_name, delimiter, files, format_str[8:], has_edge_importance)
    if format_str.startswith("udd"):
        assert data_name != "FB15k", "You should provide the dataset name for udd format."
        return KGDatasetUDD(data_path, data_name, delimiter, files, format_str[4:], has_edge_importance)
    assert False, "Unknown format {}".format(format_str)
