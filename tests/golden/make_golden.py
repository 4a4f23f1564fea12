"""Regenerate tests/golden/*.npz from the reference's committed output artefacts.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden.py
Sources:
  src/final_example.nc, src/dense_example.nc  Model204, 10 links of data/small_test.csv, stub
      forcing rain=0.001, T=1.0 (double literals), y0={0.01,3,0,5,0.2}, t in [0,2880] min,
      2881 queries, rtol 1e-6, atol 1e-9, initialStep 1e-6.  HDF5 files holding one
      deflate+shuffle chunk each (byte offsets 8252 / 32844); decoded with zlib + unshuffle
      because no HDF5/netCDF reader is installed.
  src/final.csv, src/dense.csv  DummyModel, 4 identical systems, y0=ones, t in [0,5], 10 000
      queries at (i+1)*5/10001, 6 significant digits.
  data/small_test.csv  the 10-link parameter file (copied verbatim as a fixture: it is an
      input data file of 11 lines, not source code).
  SciPy solve_ivp(RK45) on the notebook's DummyModel rhs (model_dummy_python.ipynb:65-96,150-160).
"""
import os
import shutil
import zlib

import numpy as np
import pandas as pd
from scipy.integrate import solve_ivp

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def decode_chunk(path, offset, shape):
    raw = open(path, "rb").read()
    data = zlib.decompressobj().decompress(raw[offset:])
    n = int(np.prod(shape)) * 8
    assert len(data) >= n
    # HDF5 shuffle filter: byte-plane major -> element major
    return np.frombuffer(data[:n], np.uint8).reshape(8, -1).T.copy().view("<f8").reshape(shape)


def main():
    final = decode_chunk(f"{REF}/src/final_example.nc", 8252, (10, 5))
    dense = decode_chunk(f"{REF}/src/dense_example.nc", 32844, (10, 2881, 5))
    assert np.all(final == final[0]) and np.all(dense == dense[0]), "systems expected identical"
    np.savez_compressed(f"{OUT}/model204_example.npz", final=final, dense_sys0=dense[0],
                        query_times=np.arange(2881, dtype=np.float64))

    fcsv = pd.read_csv(f"{REF}/src/final.csv").values
    dcsv = pd.read_csv(f"{REF}/src/dense.csv")
    cols = [[f"Var{i}_sys{s}" for i in range(5)] for s in range(4)]
    d0 = dcsv[cols[0]].values
    for s in range(1, 4):
        assert np.all(dcsv[cols[s]].values == d0)

    def rhs(t, y):  # notebook code cell (I2 = 0.6*H1)
        H0, H1, H2, H3, H4 = y
        Y0 = 0.5 * H0
        X2 = 0.3 * H1
        I2 = 0.6 * H1
        I3 = 0.4 * H3
        return [1.0 - Y0, 1.2 + Y0 - X2 - 0.4 - I2, X2 - 0.2, I2 - I3 - 0.3, I3 - 0.1]

    tq = (np.arange(10000) + 1) * 5.0 / 10001
    sol = solve_ivp(rhs, (0.0, 5.0), np.ones(5), method="RK45", rtol=1e-6, atol=1e-9, t_eval=tq,
                    dense_output=True)
    plain = solve_ivp(rhs, (0.0, 5.0), np.ones(5), method="RK45", rtol=1e-6, atol=1e-9)
    np.savez_compressed(f"{OUT}/dummy_example.npz", final_csv=fcsv, dense_csv_sys0=d0,
                        time_csv=dcsv["time"].values, query_times=tq,
                        scipy_final=sol.sol(5.0), scipy_dense=sol.y.T, scipy_len_t=len(plain.t),
                        scipy_nfev=plain.nfev)
    shutil.copyfile(f"{REF}/data/small_test.csv", f"{OUT}/small_test.csv")
    print("wrote", os.listdir(OUT))


if __name__ == "__main__":
    main()
