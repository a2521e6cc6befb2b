"""Golden vectors for the chunk cutting of ChunkRawRecords (strax_interface.py:368-497): the
unmodified reference class is driven by a stand-in RawData that replays a fixed list of
digitisation groups / ZLE intervals; we record the chunk bounds and the records per chunk."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def make_groups(seed, n_groups, mean_gap_s):
    rng = np.random.default_rng(seed)
    groups, t = [], 1_000_000
    for g in range(n_groups):
        t += int(rng.exponential(mean_gap_s * 1e8)) + 30_000       # samples (10 ns)
        width = int(rng.integers(400, 60_000))
        n_itv = int(rng.integers(0, 6))
        left = t - (t % 2)
        groups.append((left, left + width, n_itv))
        t += width
    return groups


def main(ref, c0_config):
    cases = {}
    # the last two: a record buffer smaller than a chunk's records (strax_interface.py:409-422 -- the class
    # allocates 5e6 records; a short buffer reaches the same branch with few groups)
    for name, seed, n_groups, gap, chunk_size, buffer_length in [
            ('sparse', 1, 40, 0.8, 2, None), ('dense', 2, 60, 0.05, 1, None), ('long_chunks', 3, 25, 1.5, 10, None),
            ('buffer_overflow_dense', 4, 80, 0.05, 1, 60), ('buffer_overflow_long', 5, 50, 0.4, 10, 36),
            # chunks much shorter than the spacing of the groups: the class closes at most one chunk per ZLE interval,
            # so its chunk clock lags behind the data
            ('lagging_clock', 6, 30, 0.4, 0.05, None), ('lagging_clock_overflow', 7, 40, 0.4, 0.05, 24)]:
        cfg, _, _ = c0_config()
        cfg['chunk_size'] = chunk_size
        groups = make_groups(seed, n_groups, gap)

        class Replay:
            def __init__(self, config):
                self.config = config
                self.source_finished = False
                self.left = self.right = 0

            def __call__(self, instructions=None, truth_buffer=None, **kw):
                for left, right, n_itv in groups:
                    self.left, self.right = left, right
                    for k in range(n_itv):
                        a = left + 60 + 200 * k
                        yield 7 + k, a, a + 119, np.full(120, 15000, np.int64)
                self.source_finished = True

        crr = ref.ChunkRawRecords(cfg, rawdata_generator=Replay)
        if buffer_length is not None:
            crr.record_buffer = crr.record_buffer[:buffer_length]
        inst = np.zeros(1, dtype=ref.strax_interface.instruction_dtype)
        inst['time'] = groups[0][0] * 10 + 600
        bounds, counts = [], []
        for res in crr(inst):
            bounds.append((int(crr.chunk_time_pre), int(crr.chunk_time)))
            counts.append(len(res['raw_records']))
        cases[name] = dict(chunk_size=chunk_size, t_min_instruction=int(inst['time'][0]), record_buffer=buffer_length,
                           groups=groups, bounds=bounds, records_per_chunk=counts)
        print(name, len(bounds), 'chunks', sum(counts), 'records')
    with open(os.path.join(HERE, 'chunks.json'), 'w') as f:
        json.dump(cases, f)
