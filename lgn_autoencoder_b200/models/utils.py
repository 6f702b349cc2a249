"""Small helpers of the model package (reference: lgn/models/utils.py:4-38)."""


def adapt_var_list(var, num_cg_levels):
    """Broadcast a scalar / short list of per-level settings to ``num_cg_levels`` entries (a longer list is cut to
    num_cg_levels - 1 entries, as the reference does)."""
    if isinstance(var, list):
        if len(var) < num_cg_levels:
            return var + (num_cg_levels - len(var)) * [var[-1]]
        if len(var) == num_cg_levels:
            return var
        return var[: num_cg_levels - 1]
    if isinstance(var, (float, int)):
        return [var] * num_cg_levels
    raise ValueError(f"Incorrect type of variables: {type(var)}. The allowed data types are list, float, or int")
