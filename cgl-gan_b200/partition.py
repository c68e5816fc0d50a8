"""Bit-exact host-side topology: dataset partitioning, client<->server assignment, FeGAN client
selection (row a11 of SURVEY.md section 8; stays in Python, integer work, negligible cost).

The reference slices tensors / torchvision datasets in place; here every function returns INDEX lists
into the caller's original dataset (same membership, same order), so the engine can keep one resident
copy of the data and gather batches by index. All functions consume the caller's `random.Random`
stream exactly as the reference consumes its global `rd` (seeded 20211212, CGLGAN/2DMG/main.py:41-42),
including the draw for `test_set` that precedes the split.

Variants (selected by `style`):
  "cgl"  CGLGAN/2DMG/main.py:382-438, CGLGAN/MNIST/main.py:386-442  (cut points in range(1, nw**2))
  "fl"   FLGAN/MNIST/flgan.py:278-334                               (cut points in range(1, nw*2))
  "fl2d" FLGAN/2DMG/flgan.py:267-322, MDGAN/2DMG/mdgan.py:289-344   (as "fl", test_set = deepcopy: no draw)
  "cap"  capgan.py:358-424, mixed-gan.py, fegan.py (torchvision datasets; iid=2 samples inside each label run)
"""
import numpy as np


def _sizes(rd, num_workers, square):
    top = num_workers ** 2 if square else num_workers * 2
    se = rd.sample(range(1, top), k=num_workers - 1)
    se.append(0)
    se.append(top)
    se = sorted(se)
    return [(se[i] - se[i - 1]) / top for i in range(1, len(se))]


def allocate_dataset(labels, iid, num_workers, num_class, num_sample, rd, style="cgl"):
    """Returns (test_idx or None, [index list per worker]). labels: 1-D integer array-like."""
    labels = np.asarray(labels)
    data_len = len(labels)
    indexes = [x for x in range(0, data_len)]
    test_idx = None
    if style != "fl2d":
        test_idx = rd.sample(range(data_len), num_sample)
    parts = []
    if iid == 0:
        sizes = [1.0 / num_workers for _ in range(num_workers)]
        rd.shuffle(indexes)
        for frac in sizes:
            part_len = int(frac * data_len)
            parts.append(indexes[0:part_len])
            indexes = indexes[part_len:]
        return test_idx, parts

    order = np.argsort(labels)          # same call (default kind) as the reference
    sorted_labels = labels[order]
    sizes = _sizes(rd, num_workers, square=style in ("cgl", "cap"))
    if iid == 1:
        lab = sorted_labels.tolist()
        for i in range(num_workers):
            index_s = (i - 1 + num_class) % num_class
            index_e = (i + 2) % num_class
            s = lab.index(index_s)
            e = lab.index(index_e)
            l = int(sizes[i] * data_len)
            if s < e:
                if l > (e - s):
                    l = e - s
                pick = rd.sample(range(s, e), l)
            else:
                if l > (e + data_len - s):
                    l = e + data_len - s
                pick = rd.sample(list(range(0, e)) + list(range(s, data_len)), l)
            parts.append([int(order[j]) for j in pick])
        return test_idx, parts

    if style == "cap":
        # capgan.py:412-424: a sampled subset of each label run, wrapping around
        l, s = 1, 0
        for i in range(num_workers):
            while l < data_len and sorted_labels[l] == sorted_labels[l - 1]:
                l += 1
            pick = rd.sample(range(s, l), min(int(sizes[i] * data_len), l - s))
            parts.append([int(order[j]) for j in pick])
            s = l % data_len
            l = s + 1
        return test_idx, parts

    # tensor form (CGLGAN/2DMG/main.py:430-438): peel one label run per worker off the front; the scan
    # stops at len-1, so the very last sample is never handed out.
    start = 0
    remaining = data_len
    for i in range(num_workers):
        l = 1
        while sorted_labels[start + l] == sorted_labels[start + l - 1] and l < remaining - 1:
            l += 1
        parts.append([int(order[j]) for j in range(start, start + l)])
        start += l
        remaining -= l
    return test_idx, parts


def assign_clients(num_workers, num_servers):
    """Contiguous block assignment (CGLGAN/2DMG/main.py:468-474): returns (client_list per server,
    server_list per worker). Workers beyond num_servers * (num_workers // num_servers) get no server."""
    worker = [i for i in range(num_workers)]
    client_list = [[] for _ in range(num_servers)]
    server_list = [[] for _ in range(num_workers)]
    k = num_workers // num_servers
    for i in range(num_servers):
        al = worker[:k]
        worker = worker[k:]
        for j in al:
            client_list[i].append(j)
            server_list[j].append(i)
    return client_list, server_list


def init_groups(size, cls_freq_wrk, frac_workers, max_groups=20000, num_class=10):
    """FeGAN balanced client selection (fegan.py:383-452): every group greedily takes, for the class
    least represented so far, the next worker in that class's round-robin queue."""
    from collections import deque
    gp_size = max(1, int(frac_workers * size))
    wrk_cls = [[freq != 0 for freq in cls_list] for cls_list in cls_freq_wrk]
    cls_q = [deque() for _ in range(num_class)]
    for worker, class_list in enumerate(reversed(wrk_cls)):
        for cls, exist in enumerate(class_list):
            if exist:
                cls_q[cls].append(size - worker - 1)
    taken_count = np.zeros(num_class, dtype=np.asarray(cls_freq_wrk[0]).dtype)
    groups, choose_r = [], []
    while True:
        visited = [False for _ in range(size)]
        g = []
        for _ in range(gp_size):
            cls = int(np.where(taken_count == np.amin(taken_count))[0][0])
            done_q = False
            count = 0
            while not done_q:
                wrkr = cls_q[cls].popleft()
                if not visited[wrkr] and wrk_cls[wrkr][cls]:
                    g.append(wrkr)
                    taken_count = taken_count + np.asarray(cls_freq_wrk[wrkr])
                    visited[wrkr] = True
                    done_q = True
                cls_q[cls].append(wrkr)
                count += 1
                if count == size:
                    done_q = True
        choose_r.append(0 in g)
        groups.append(g)
        if len(groups) >= max_groups:
            break
    return groups, choose_r


def gmm_labels_and_data(n_class, x, seed=20211212):
    """The synthetic 2-D Gaussian-mixture ring (CGLGAN/2DMG/data.py:23-38), vectorised: same mixture
    (radius 1, sigma 0.01, x*n_class points, classes drawn by np.random.randint after np.random.seed(seed)),
    sorted by label. Class draws are bit-identical to the reference; the Gaussian noise comes from numpy
    instead of one torch.normal call per point, so coordinates match in distribution only."""
    import torch
    rs = np.random.RandomState(seed)
    thetas = np.linspace(0, 2 * (1 - 1 / n_class) * np.pi, n_class)
    xs, ys = np.sin(thetas), np.cos(thetas)
    n = x * n_class
    coins = np.array([rs.randint(0, n_class) for _ in range(n)])
    noise = np.random.RandomState(seed + 1).normal(0.0, 0.01, size=(n, 2))
    data = np.stack([xs[coins], ys[coins]], axis=1) + noise
    labels = torch.from_numpy(coins).float()
    targets, indexes = torch.sort(labels)
    return torch.from_numpy(data).float()[indexes], targets
