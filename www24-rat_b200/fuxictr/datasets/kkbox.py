"""KKBox feature engineering hooks (reference fuxictr/datasets/kkbox.py:23-49), named in the dataset YAML as `preprocess:`."""
import numpy as np
import pandas as pd

from ..features import FeatureEncoder as BaseFeatureEncoder


class FeatureEncoder(BaseFeatureEncoder):
    def extract_country_code(self, df, col_name):
        """ISRC -> its two-letter country prefix ('' for a missing code)"""
        s = df[col_name]
        return s.str.slice(0, 2).where(s.notnull(), "")

    def bucketize_age(self, df, col_name):
        """age -> decade bucket '1'..'7' ((0,10], (10,20], ... , > 60); '' for missing or implausible (< 1 or > 95) ages"""
        age = pd.to_numeric(df[col_name], errors="coerce").astype(float).values
        bucket = np.minimum(np.ceil(age / 10.0), 7.0)
        ok = ~np.isnan(age) & (age >= 1) & (age <= 95)
        out = np.full(len(age), "", dtype=object)
        out[ok] = bucket[ok].astype(int).astype(str)
        return pd.Series(out, index=df.index)
