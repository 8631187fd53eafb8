"""ann-benchmarks HDF5 (train / test / distances datasets, benchmark/dataset.ml:76-102) -> .fbin files, through the
package's own HDF5 reader (ocaml-hnsw_b200/hdf5min.py; no h5py needed).  usage: hdf5_to_fbin.py file.hdf5 out_prefix"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ocaml_hnsw_b200 as H
from ocaml_hnsw_b200.hdf5min import Hdf5File


def main(src, prefix):
    with Hdf5File(src) as f:
        for name in ("train", "test", "distances"):
            a = np.ascontiguousarray(f[name], np.float32)
            H.write_fbin(f"{prefix}.{name}.fbin", a)
            print(name, a.shape)
        print("distance:", f.attrs.get("distance"))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
