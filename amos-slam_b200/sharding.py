"""Frame / sequence sharding across the GPUs of one box: the path has no cross-frame state, so units are
partitioned with no data-path collective.  unit = frame (C1, C3), frame pair (C2), stereo pair on one GPU (C4),
sequence-affine round-robin for camera streams (C5: gpu = seq_id mod G, one extractor handle + CUDA stream per stream).
The sequence -> (gpu, stream) rule is the product's own: orbx_pool_shard_of of liborbx_b200.so (the rule orbx_pool applies)."""
import ctypes as C


def shard_range(n_units, rank, world):
    """Contiguous, balanced [begin, end) of `n_units` independent units owned by `rank` of `world` (bench.py under torchrun)."""
    if world <= 0 or not (0 <= rank < world) or n_units < 0:
        raise ValueError("bad shard arguments")
    base, rem = divmod(n_units, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def worker_for_sequence(seq_id, n_gpus, streams_per_gpu=1):
    """(gpu, stream) that serves a camera stream: it stays on one worker (its extractor handle is stateful: mvImagePyramid)."""
    from . import lib
    g, s = C.c_int(), C.c_int()
    if lib().orbx_pool_shard_of(int(seq_id), int(n_gpus), int(streams_per_gpu), C.byref(g), C.byref(s)) != 0:
        raise ValueError("bad sequence / pool shape")
    return g.value, s.value


def gpu_for_sequence(seq_id, n_gpus):
    return worker_for_sequence(seq_id, n_gpus)[0]


def sequences_of_gpu(n_sequences, gpu, n_gpus):
    return [s for s in range(n_sequences) if gpu_for_sequence(s, n_gpus) == gpu]
