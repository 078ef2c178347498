"""TEST INFRASTRUCTURE ONLY -- freezes /root/reference/data/sharkTrackingData.csv (the raw shark tracks
BASELINE.json's config 3 names) as tests/golden/shark_tracks_raw.npz, because /root/reference does not
exist on the GPU box.  Layout measured in the survey: 32 sharks x 4 rows (x, vx, y, vy) x 815 frames;
only x / y are kept (raw pixel-like units, as float64)."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
a = np.loadtxt("/root/reference/data/sharkTrackingData.csv", delimiter=",")
x, y = a[0::4], a[2::4]
assert x.shape == (32, 815)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "shark_tracks_raw.npz"), x=x, y=y)
print("wrote", x.shape)
