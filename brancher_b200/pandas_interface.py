"""DataFrame <-> sample-dict conversions used by the user-facing `get_sample` API
(mirror of brancher/pandas_interface.py; presentation only, not on the hot path)."""
from collections.abc import Iterable

import numpy as np
import pandas as pd

from brancher_b200.utilities import is_tensor, to_numpy


def pandas_frame2value(dataframe, index):
    if isinstance(dataframe, pd.DataFrame):
        return np.array([x.tolist() if isinstance(x, np.ndarray) else x for x in dataframe[index].values])
    return dataframe


def _reformat_value(value, index):
    if is_tensor(value):
        row = to_numpy(value[index])
        if row.size == 1:
            return float(row.reshape(()))
        if value.shape[1] == 1:
            return row[0]
        return to_numpy(value)
    if isinstance(value, dict):
        return {k: _reformat_value(v, index) for k, v in value.items()}
    if isinstance(value, Iterable) and not isinstance(value, str):
        return [_reformat_value(v, index) for v in value]
    return value


def reformat_sample_to_pandas(sample):
    tensors = [v for v in sample.values() if is_tensor(v)]
    n = max(t.shape[0] for t in tensors) if tensors else 1
    data = [[_reformat_value(value, i if (not is_tensor(value) or value.shape[0] > 1) else 0) for i in range(n)]
            for value in sample.values()]
    return pd.DataFrame(data, index=[k.name for k in sample.keys()], columns=range(n)).transpose()


def reformat_model_summary(summary_data, var_names, feature_list):
    return pd.DataFrame(summary_data, index=var_names, columns=feature_list).transpose()
