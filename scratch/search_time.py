import time, numpy as np, torch, sys
sys.path.insert(0, '.')
import triple_accel_b200 as ta
from triple_accel_b200 import synth
eng = ta.Engine(0)
needle, hay, hoff = synth.needle_haystacks(100000, 4096, 32, 0.01, 3, seed=1234)
d_hay = torch.from_numpy(hay).cuda(); d_off = torch.from_numpy(hoff.view(np.int64)).cuda()
for _ in range(3): eng.levenshtein_search_batch_dev(needle, d_hay, d_off, 4096, 3, 1)
import cProfile, pstats
t0 = time.perf_counter()
for _ in range(20): r = eng.levenshtein_search_batch_dev(needle, d_hay, d_off, 4096, 3, 1)
print("per call ms", (time.perf_counter() - t0) / 20 * 1e3)
pr = cProfile.Profile(); pr.enable()
for _ in range(20): r = eng.levenshtein_search_batch_dev(needle, d_hay, d_off, 4096, 3, 1)
pr.disable(); pstats.Stats(pr).sort_stats('cumulative').print_stats(8)
