"""Down-sampled slice of the reference's recorded path ``example_paths/twoRooms_path.npy`` (60 000 x 2 float64, 960 KB) as
a small fixture: every 20th sample, float32 (3 000 x 2, 24 KB).  Loaded with ``data_dt = 20 * dt`` it is re-sampled by the
drivers' own ``stretch_trajectory`` rule (``run_slam.py:80-89,100-104``) back to 60 000 steps of 1 ms.

    python tests/golden/make_path_fixture.py            (authoring container: /root/reference mounted)
"""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/example_paths/twoRooms_path.npy"


def main():
    path = np.load(SRC)
    assert path.shape == (60000, 2)
    np.save(os.path.join(HERE, "twoRooms_path_ds20.npy"), path[::20].astype(np.float32))
    print("wrote", os.path.join(HERE, "twoRooms_path_ds20.npy"), path[::20].shape)


if __name__ == "__main__":
    main()
