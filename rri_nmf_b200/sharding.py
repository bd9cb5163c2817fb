"""Row partition of a multi-GPU run: rank r owns rows [b_r, e_r) of X and W; T is replicated.
(The reference defines the shard statistic at nmf.py:680-686; it has no multi-process code.)"""


def shard_bounds(n, world):
    """balanced contiguous row ranges, the first n % world shards one row longer"""
    q, r = divmod(int(n), int(world))
    out, b = [], 0
    for p in range(world):
        e = b + q + (1 if p < r else 0)
        out.append((b, e))
        b = e
    return out
