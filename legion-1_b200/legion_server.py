#!/usr/bin/env python
"""Launcher with the reference's command line (legion_server.py:72-85): picks the dataset shape, writes the
11-field ./meta_config (GPUGraphStore.cu:190-223) and starts the B200 server binary with <gpu_number>
<cache_agg_mode>.  Differences from the reference launcher: the binary is legion-1_b200/_build/legion, custom
datasets can be described with --vertices/--edges/--features_dim/..., and --cache_agg_mode can be forced (the
reference only ever selects GPUs-per-clique 1 or 2, legion_server.py:62-68)."""
import argparse
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
LEGION = os.path.join(HERE, "_build", "legion")

# name -> (subdir, vertices, edges, feature dim, train, valid, test)   (legion_server.py:6-53)
DATASETS = {
    "PR": ("products", 2449029, 123718280, 100, 196615, 39323, 2213091),
    "PA": ("paper100M", 111059956, 1615685872, 128, 11105995, 100000, 100000),
    "CO": ("com-friendster", 65608366, 1806067135, 256, 6560836, 100000, 100000),
    "UKS": ("ukunion", 133633040, 5507679822, 256, 13363304, 100000, 100000),
    "UKL": ("uk2014", 787801471, 47284178505, 128, 78780147, 100000, 100000),
    "CL": ("clueweb", 955207488, 42574107469, 128, 95520748, 100000, 100000),
}


def main(argv=None):
    ap = argparse.ArgumentParser("Legion Server (B200).")
    ap.add_argument("--dataset_path", type=str, default="/home/atc-artifacts-user/datasets")
    ap.add_argument("--dataset", type=str, default="PA")
    ap.add_argument("--train_batch_size", type=int, default=8000)
    ap.add_argument("--hops_num", type=int, default=2)
    ap.add_argument("--nbrs_num", type=str, default="25,10")
    ap.add_argument("--gpu_number", type=int, default=1)
    ap.add_argument("--epoch", type=int, default=10)
    ap.add_argument("--cache_memory", type=int, default=38000000000)
    ap.add_argument("--usenvlink", type=int, default=1)
    ap.add_argument("--cache_agg_mode", type=int, default=-1, help="-1: one clique over all GPUs (NVSwitch); 0/1/2/3 = 1/2/4/8 GPUs per clique")
    ap.add_argument("--rng", type=str, default="minstd", choices=["minstd", "philox"])
    ap.add_argument("--custom", type=str, default="", help="path,vertices,edges,dim,train,valid,test for a dataset outside the table")
    ap.add_argument("--dry_run", action="store_true")
    a = ap.parse_args(argv)
    if a.custom:
        f = a.custom.split(",")
        path, shape = f[0], tuple(int(x) for x in f[1:7])
    elif a.dataset in DATASETS:
        sub, *shape = DATASETS[a.dataset]
        path = os.path.join(a.dataset_path, sub)
    else:
        print("invalid dataset path")
        return 2
    if not path.endswith("/"):
        path += "/"
    v, e, d, ntr, nva, nte = shape
    with open("meta_config", "w") as fh:
        fh.write("{} {} {} {} {} {} {} {} {} {} {}".format(path, a.train_batch_size, v, e, d, ntr, nva, nte, a.cache_memory, a.epoch,
                                                          1 - a.usenvlink))
    mode = a.cache_agg_mode
    if mode < 0:
        # The reference pairs GPUs (Kg <= 2, legion_server.py:62-68) because its testbeds had NVLink bridges between pairs.
        # On an NVSwitch node every GPU reaches every peer at full bandwidth, so the clique is the whole machine: Kg = P
        # (mode 0/1/2/3 = 1/2/4/8 GPUs per clique, GPUCache.cu:593-607); a GPU count that is not a power of two keeps pairs.
        mode = {1: 0, 2: 1, 4: 2, 8: 3}.get(a.gpu_number, 1 if a.gpu_number >= 2 else 0) if a.usenvlink == 1 else 0
    env = dict(os.environ, LEGION_FANOUT=a.nbrs_num.strip("[] ").replace(" ", ""), LEGION_RNG=a.rng)
    cmd = [LEGION, str(a.gpu_number), str(mode)]
    if a.dry_run:
        print(" ".join(cmd))
        return 0
    os.execve(LEGION, cmd, env)


if __name__ == "__main__":
    sys.exit(main())
