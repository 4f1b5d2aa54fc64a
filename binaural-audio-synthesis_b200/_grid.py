"""The 187-point measurement grid (sphere.py:124-319) as pure numpy: shared by sphere.py and by
bank_synth.py, which must stay importable without libbas_b200.so (bench.py's reference arm builds the
synthetic bank without loading the product library)."""
import numpy as np

RING_ELEV_DEG = (-45, -30, -15, 0, 15, 30, 45, 60, 75, 90)
RING_COUNT = (24, 24, 24, 24, 24, 24, 24, 12, 6, 1)


def get_index_elev_azim() -> np.ndarray:
    """sphere.py:124-319."""
    rows = []
    index = 0
    for elev, count in zip(RING_ELEV_DEG, RING_COUNT):
        for k in range(count):
            rows.append((index, elev, k * (360 // count) if count > 1 else 0))
            index += 1
    table = np.array(rows, dtype=np.float32)
    table[:, 1:3] *= (2 * np.pi / 360)          # float32 multiply, like sphere.py:318
    return table
