"""ann-benchmarks HDF5 (train / test / distances datasets, benchmark/dataset.ml:76-102) -> .fbin files.
Needs h5py, which this image does not have; run it wherever the data was downloaded."""
import sys
import numpy as np

def main(src, prefix):
    import h5py
    with h5py.File(src, "r") as f:
        for name in ("train", "test", "distances"):
            a = np.ascontiguousarray(f[name][...], np.float32)
            with open(f"{prefix}.{name}.fbin", "wb") as out:
                np.array(a.shape, np.int32).tofile(out)
                a.tofile(out)
            print(name, a.shape)
        print("distance:", f.attrs.get("distance"))

if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
