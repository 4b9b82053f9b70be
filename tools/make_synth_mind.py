"""Tiny synthetic stand-in for the reference's PARSED MIND files (the output of its data_preprocess.py), enough to drive
the reference `train.py` end to end:  data/train/behaviors_parsed.tsv  (user, clicked_news, candidate_news, clicked) and
data/train/news_parsed.tsv  (id, category, subcategory, title, abstract, title_entities, abstract_entities; list columns
as Python literals) -- formats from reference src/dataset.py:26-85 and src/data_preprocess.py:77-81,150-154.

    python tools/make_synth_mind.py OUT_DIR [n_behaviors] [n_news] [K]
"""
import os
import random
import sys

out = sys.argv[1]
n_beh = int(sys.argv[2]) if len(sys.argv) > 2 else 64
n_news = int(sys.argv[3]) if len(sys.argv) > 3 else 200
K = int(sys.argv[4]) if len(sys.argv) > 4 else 2  # negative_sampling_ratio of the reference config
rng = random.Random(0)
os.makedirs(os.path.join(out, "data", "train"), exist_ok=True)
T, TA = 20, 50


def padded(n, length):
    ids = [rng.randint(1, 999) for _ in range(n)]
    return ids + [0] * (length - n)


with open(os.path.join(out, "data", "train", "news_parsed.tsv"), "w") as f:
    f.write("id\tcategory\tsubcategory\ttitle\tabstract\ttitle_entities\tabstract_entities\n")
    for i in range(n_news):
        f.write(f"N{i}\t{rng.randint(1, 17)}\t{rng.randint(1, 200)}\t{padded(rng.randint(5, T), T)}\t{padded(rng.randint(10, TA), TA)}\t"
                f"{[0] * T}\t{[0] * TA}\n")
with open(os.path.join(out, "data", "train", "behaviors_parsed.tsv"), "w") as f:
    f.write("user\tclicked_news\tcandidate_news\tclicked\n")
    for b in range(n_beh):
        hist = " ".join(f"N{rng.randrange(n_news)}" for _ in range(rng.randint(1, 60)))
        cand = " ".join(f"N{rng.randrange(n_news)}" for _ in range(1 + K))
        f.write(f"{rng.randint(1, 400)}\t{hist}\t{cand}\t{' '.join(['1'] + ['0'] * K)}\n")
print("wrote", out)
