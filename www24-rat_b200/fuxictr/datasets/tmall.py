"""Tmall feature engineering hooks (reference fuxictr/datasets/tmall.py:25-37): weekday / weekend of the `time_stamp` column
("MDD" / "MMDD" strings of the year 2014)."""
import pandas as pd

from ..features import FeatureEncoder as BaseFeatureEncoder


def _weekday(series):
    """'%w' weekday ('0' = Sunday .. '6' = Saturday) of month/day strings in 2014"""
    ts = series.astype(str)
    dates = pd.to_datetime(dict(year=2014, month=ts.str.slice(0, -2).astype(int), day=ts.str.slice(-2).astype(int)))
    return ((dates.dt.dayofweek + 1) % 7).astype(str)        # pandas: Monday = 0 ; strftime('%w'): Sunday = 0


class FeatureEncoder(BaseFeatureEncoder):
    def convert_weekday(self, df, col_name):
        return _weekday(df["time_stamp"])

    def convert_weekend(self, df, col_name):
        return _weekday(df["time_stamp"]).isin(["6", "0"]).map({True: "1", False: "0"})
