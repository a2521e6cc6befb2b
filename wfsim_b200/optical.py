"""Host-side preparation of externally supplied (optical Monte-Carlo) photons: the step beside the hot path that
turns per-entry photon lists into instructions the simulator takes.

Mirror of wfsim/utils.py:61-165 (`find_optical_t_range`, `split_long_optical_pulse`, `optical_adjustment`), vectorised
where the reference loops in numba, with its behaviour kept to the byte -- including that entries split off a long
pulse keep the time of their parent and photon times relative to the parent's first photon (the reference's second
pass starts behind the appended rows and therefore never touches them).  Reading the G4 files themselves
(`read_optical`, strax_interface.py:286-333) needs uproot, a third-party package that is not in this image; its
output -- (instructions with `_first` / `_last`, channels, timings) -- is what these functions and
`Simulator.simulate(..., optical=...)` take."""
import numpy as np

PULSE_MAX_DURATION = int(1e3)      # utils.py:9
N_SPLIT_LOOP = 5                   # utils.py:10


def find_optical_t_range(firsts, lasts, timings, tmins, tmaxs, start=0):
    """utils.py:61-87: min / max photon time of every entry from `start` on (-1 / -1 for an empty one); the entry's
    photon times become relative to its first photon.  In place."""
    firsts, lasts = np.asarray(firsts)[start:], np.asarray(lasts)[start:]
    n = lasts - firsts
    empty = n == 0
    tmins[start:][empty] = -1
    tmaxs[start:][empty] = -1
    rows = np.flatnonzero(~empty)
    if not len(rows):
        return
    # photon index lists of the non-empty entries, entry by entry (entries need not tile the photon arrays)
    reps = n[rows]
    entry = np.repeat(np.arange(len(rows)), reps)
    offs = np.arange(reps.sum()) - np.repeat(np.cumsum(reps) - reps, reps)
    idx = firsts[rows][entry] + offs
    t = timings[idx]
    lo = np.full(len(rows), np.iinfo(np.int64).max, np.int64)
    hi = np.full(len(rows), np.iinfo(np.int64).min, np.int64)
    np.minimum.at(lo, entry, t)
    np.maximum.at(hi, entry, t)
    tmins[start:][rows] = lo
    tmaxs[start:][rows] = hi
    # (overlapping entries would be shifted once per entry in the reference too; subtract.at keeps that)
    np.subtract.at(timings, idx, lo[entry])


def split_long_optical_pulse(firsts, lasts, timings, channels):
    """utils.py:90-118: the photons of an entry that arrive more than PULSE_MAX_DURATION ns after its first photon
    are swapped to the front of the entry (in the reference's order of swaps) and handed out as an entry of their
    own: yields (index, first, last) and leaves firsts[index] behind the split-off photons."""
    for ix in range(len(firsts)):
        f, l = int(firsts[ix]), int(lasts[ix])
        late = f + np.flatnonzero(timings[f:l] > PULSE_MAX_DURATION)
        if len(late) == 0:
            continue
        cnt = f
        for k, iy in enumerate(late):
            cnt = k + f
            if iy > cnt:
                timings[cnt], timings[iy] = timings[iy], timings[cnt]
                channels[cnt], channels[iy] = channels[iy], channels[cnt]
        yield ix, f, cnt + 1
        firsts[ix] = cnt + 1


def optical_adjustment(instructions, timings, channels):
    """utils.py:121-165: (1) the instruction time moves to the first photon of its entry, photon times become
    relative to it; (2) photons later than PULSE_MAX_DURATION are split off into new instructions appended at the end.
    `timings` / `channels` are changed in place; the (possibly longer) instruction array is returned."""
    tmins = np.zeros(len(instructions), np.int64)
    tmaxs = np.zeros(len(instructions), np.int64)
    start = 0
    for _ in range(N_SPLIT_LOOP):
        find_optical_t_range(instructions['_first'], instructions['_last'], timings, tmins, tmaxs, start=start)
        instructions['time'][start:] += tmins[start:]
        long_pulse = ((tmaxs - tmins) > PULSE_MAX_DURATION) & (np.arange(len(instructions)) >= start)
        if long_pulse.sum() < 1:
            break
        where = np.flatnonzero(long_pulse)
        extra = []
        for ix, first, last in split_long_optical_pulse(instructions['_first'][long_pulse], instructions['_last'][long_pulse],
                                                        timings, channels):
            tmp = instructions[where[ix]].copy()
            tmp['_first'], tmp['_last'] = first, last
            instructions['_first'][where[ix]] = last
            extra.append(tmp)
        instructions = np.append(instructions, np.array(extra, dtype=instructions.dtype))
        tmins = np.hstack([tmins, np.zeros(len(extra), np.int64)])
        tmaxs = np.hstack([tmaxs, np.zeros(len(extra), np.int64)])
        start = len(instructions)
    return instructions
