"""plot_2d's KL score, restated (CGLGAN/2DMG/main.py:68-94): numpy histogram2d + scipy entropy, as the reference.
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py)."""
import numpy as np
from scipy.stats import entropy


def kl_score_2d(real_points, generated_points, stepsize=16):
    """real_points / generated_points: [n, 2] arrays (the reference's `sd` and `D`, main.py:68,83)."""
    r, g = np.asarray(real_points), np.asarray(generated_points)
    count_r, _, _ = np.histogram2d(r[:, 0], r[:, 1], bins=stepsize, range=[[-1, 1], [-1, 1]])   # main.py:72
    count_g, _, _ = np.histogram2d(g[:, 0], g[:, 1], bins=stepsize, range=[[-1, 1], [-1, 1]])   # main.py:85
    r_h, g_h = [], []
    for i in range(len(count_r)):                                                               # main.py:86-91
        for j in range(len(count_r)):
            if count_r[i][j] != 0:
                r_h.append(count_r[i][j])
                g_h.append(count_g[i][j])
    return entropy(g_h, r_h)                                                                    # main.py:94
