"""Deterministic synthetic inputs for the unmodified reference drivers (no dataset is available offline).

text: a 5-topic template language (topic decides which fifth of the 200-word content vocabulary a sentence draws from;
the first 25 training sentences enumerate the vocabulary so that V = 209 >= 128 and the tcgen05 path is the one that
runs), one sentence per line, lengths 4..11 tokens — MonoTextData (data/text_data.py:65-279) builds the vocabulary from the train file
and buckets equal-length sentences into batches, including ragged ones.
image: random smooth blobs in [0,1] as (x_train, x_val, x_test), the tuple image.py:204-205 loads with torch.load."""
import os

import numpy as np


def make_text(root, n_train=160, n_val=48, n_test=48, seed=7):
    d = os.path.join(root, "datasets", "tinysyn_data")
    os.makedirs(d, exist_ok=True)
    rng = np.random.RandomState(seed)
    content = ["w%d" % i for i in range(200)]
    func = ["the", "of", "and", "to", "in"]

    def sent():
        topic = rng.randint(5)
        n = rng.randint(4, 12)
        words = []
        for j in range(n):
            if j % 3 == 2:
                words.append(func[rng.randint(len(func))])
            else:
                words.append(content[topic * 40 + min(39, int(rng.exponential(8.0)))])
        return " ".join(words)

    for name, n in (("train", n_train), ("valid", n_val), ("test", n_test)):
        with open(os.path.join(d, name + ".txt"), "w") as f:
            k = 0
            if name == "train":
                for k in range(25):
                    f.write(" ".join(content[8 * k: 8 * k + 8]) + "\n")
                k = 25
            for _ in range(n - k):
                f.write(sent() + "\n")
    return d


def make_image(root, n_train=250, n_val=50, n_test=500, seed=11):   # image.py:175 needs >= 10 test batches of 50
    import torch
    d = os.path.join(root, "datasets", "tinyomni_data")
    os.makedirs(d, exist_ok=True)
    g = torch.Generator().manual_seed(seed)

    def blobs(n):
        yy, xx = torch.meshgrid(torch.arange(28.0), torch.arange(28.0), indexing="ij")
        out = torch.zeros(n, 1, 28, 28)
        for i in range(n):
            for _ in range(3):
                cx, cy = (torch.rand(2, generator=g) * 20 + 4).tolist()
                s = float(torch.rand(1, generator=g)) * 3 + 1.5
                out[i, 0] += torch.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / (2 * s * s))
        return out.clamp_(0, 1)

    torch.save((blobs(n_train), blobs(n_val), blobs(n_test)), os.path.join(d, "tinyomni.pt"))
    return d
