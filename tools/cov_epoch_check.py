"""The owner map of computeCovariance is cleared only every 509 batches (batch-epoch tags in the claims): run across two
clears, check bit-exactness against the oracle AND that no batch falls back to the sequential replay."""
import sys


def main():
    import numpy as np
    sys.path.insert(0, '.')
    from oracle import sp_oracle as O
    from sp_orb_slam_b200 import SPExtractor, synth
    H, W = 240, 320
    ex = SPExtractor(800, H, W, 'tests/golden/superpoint_v1.spw', max_batch=2)
    fr = synth.make_stream(H, W, 6, seed=23, n_shapes=300)
    bad = replayed = 0
    for it in range(1100):
        o = ex.extract_batch([fr[it % 6], fr[(it + 1) % 6]])
        replayed += int(ex.debug_read(0, "cov_replayed", 2)[:, 0].sum())
        if it % 97 == 0 or it in (0, 1, 2, 508, 509, 510, 1017, 1018, 1019):
            for x in o:
                r, c2, c2i = O.covariance(x["heat_inv"], x["kp_xy"])
                bad += not (np.array_equal(x["cov2"], c2) and np.array_equal(x["kp_response"], r))
    print("epoch test: mismatching frames", bad, " keypoints sent to the sequential replay over 1100 batches:", replayed)
    ex.close()


if __name__ == '__main__':
    main()
