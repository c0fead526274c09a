"""On-disk dataset format of the reference loader (GPUGraphStore.cu:254-301, legion_server.py:55-56):
a directory with raw little-endian arrays `edge_src` (int64 indptr, N+1), `edge_dst` (int32 indices, E),
`features` (float32 N x D), `labels` (int32 N), `trainingset` / `validationset` / `testingset` (int32 ids)
and a one-line `meta_config` in the server's working directory."""
import os

import numpy as np


def write_dataset(path, ds):
    os.makedirs(path, exist_ok=True)
    np.asarray(ds.indptr, np.int64).tofile(os.path.join(path, "edge_src"))
    np.asarray(ds.indices, np.int32).tofile(os.path.join(path, "edge_dst"))
    np.asarray(ds.features, np.float32).tofile(os.path.join(path, "features"))
    np.asarray(ds.labels, np.int32).tofile(os.path.join(path, "labels"))
    np.asarray(ds.train_ids, np.int32).tofile(os.path.join(path, "trainingset"))
    np.asarray(ds.valid_ids, np.int32).tofile(os.path.join(path, "validationset"))
    np.asarray(ds.test_ids, np.int32).tofile(os.path.join(path, "testingset"))


def write_meta_config(workdir, dataset_path, ds, batch, cache_memory, epochs, partition_flag=0):
    """11 whitespace-separated fields (GPUGraphStore::ReadMetaFIle, GPUGraphStore.cu:190-223)."""
    if not dataset_path.endswith("/"):
        dataset_path += "/"
    line = "{} {} {} {} {} {} {} {} {} {} {}".format(dataset_path, batch, ds.n_nodes, ds.n_edges, ds.dim, len(ds.train_ids),
                                                     len(ds.valid_ids), len(ds.test_ids), cache_memory, epochs, partition_flag)
    with open(os.path.join(workdir, "meta_config"), "w") as f:
        f.write(line)
    return line


def read_dataset(path, n_nodes, n_edges, dim, n_train, n_valid, n_test):
    from .synth import Dataset
    rd = lambda name, dt, n: np.fromfile(os.path.join(path, name), dtype=dt, count=n)
    return Dataset(n_nodes=n_nodes, n_edges=n_edges, dim=dim, indptr=rd("edge_src", np.int64, n_nodes + 1),
                   indices=rd("edge_dst", np.int32, n_edges), features=rd("features", np.float32, n_nodes * dim).reshape(n_nodes, dim),
                   labels=rd("labels", np.int32, n_nodes), train_ids=rd("trainingset", np.int32, n_train),
                   valid_ids=rd("validationset", np.int32, n_valid), test_ids=rd("testingset", np.int32, n_test))
