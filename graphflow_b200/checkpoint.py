"""Checkpoint files in the reference's own format, so parameters move between the two implementations.

`save_model` / `load_model` of every reference model (e.g. GraphFlow/SMP_beta.h:980-1002) dump
`sgd->params[i]->value[j]` in registration order (H, K_l, b_l ..., W) with `file << value << " "`: plain text,
whitespace separated, the stream's default formatting (6 significant digits, the `%g` rules), one trailing blank, no
optimizer state.  `load_model` reads them back with `file >> value`."""
import numpy as np


def format_value(v):
    """What `std::ostream << double` prints with default flags and precision 6."""
    v = float(v)
    if v != v:
        return "-nan" if np.signbit(v) else "nan"
    if v in (float("inf"), float("-inf")):
        return "inf" if v > 0 else "-inf"
    return "%g" % v


def save_model(path, flat_params):
    """flat_params: the flat parameter vector (numpy array or torch tensor) in registration order."""
    values = flat_params.detach().cpu().numpy() if hasattr(flat_params, "detach") else np.asarray(flat_params)
    with open(path, "w") as f:
        f.write("".join(format_value(v) + " " for v in values.ravel()))


def load_model(path, count=None, dtype=np.float32):
    """Returns the values as a flat array; `count` (if given) must match what the file holds."""
    with open(path) as f:
        values = np.array([float(tok) for tok in f.read().split()], dtype=dtype)
    if count is not None and values.size != count:
        raise ValueError("checkpoint %s holds %d values, the model has %d parameters" % (path, values.size, count))
    return values
