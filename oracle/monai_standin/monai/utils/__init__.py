def ensure_tuple_rep(tup, dim):
    """monai.utils.ensure_tuple_rep: scalar -> repeated tuple; tuple of the right length is returned as a tuple."""
    if isinstance(tup, (str, bytes)) or not hasattr(tup, "__iter__"):
        return (tup,) * dim
    tup = tuple(tup)
    if len(tup) == dim:
        return tup
    raise ValueError(f"Sequence must have length {dim}, got {len(tup)}.")
