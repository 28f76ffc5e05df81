"""Frame / sequence sharding across the GPUs of one box: the path has no cross-frame state, so units are
partitioned with no data-path collective.  unit = frame (C1, C3), frame pair (C2), stereo pair on one GPU (C4),
sequence-affine round-robin for camera streams (C5: gpu = seq_id mod G, one extractor handle + CUDA stream per stream)."""


def shard_range(n_units, rank, world):
    """Contiguous, balanced [begin, end) of `n_units` independent units owned by `rank` of `world`."""
    if world <= 0 or not (0 <= rank < world) or n_units < 0:
        raise ValueError("bad shard arguments")
    base, rem = divmod(n_units, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def gpu_for_sequence(seq_id, n_gpus):
    """A camera stream stays on one GPU (its extractor handle is stateful: mvImagePyramid)."""
    if n_gpus <= 0:
        raise ValueError("n_gpus must be positive")
    return seq_id % n_gpus


def sequences_of_gpu(n_sequences, gpu, n_gpus):
    return [s for s in range(n_sequences) if gpu_for_sequence(s, n_gpus) == gpu]
