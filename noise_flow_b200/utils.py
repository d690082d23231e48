"""Host-side utilities of the reference's training driver that sit next to the hot path: the TSV result logger
(``borealisflows/utils.py:89-107``) and the column sets ``train_noise_flow.py:334-348`` writes with it."""
from typing import Dict, Iterable, List

# train_noise_flow.py:336-348: 'NLL' = negative log-likelihood, 'NLL_G' / 'NLL_SDN' = Gaussian / camera-NLF baselines
# (metrics.calc_baselines), 'sdz' = std. dev. of the base measure, 'KLD_*' = marginal KL divergences (metrics.kl_div_3_data)
LOG_COLUMNS = ["epoch", "NLL", "NLL_G", "NLL_SDN", "sdz"]
TRAIN_COLUMNS = LOG_COLUMNS + ["train_time"]
TEST_COLUMNS = LOG_COLUMNS + ["msg"]
SAMPLE_COLUMNS = LOG_COLUMNS + ["KLD_G", "KLD_NLF", "KLD_NF", "KLD_R"]
SAMPLE_COLUMNS_DO_SAMPLE = LOG_COLUMNS + ["sample_time", "KLD_G", "KLD_NLF", "KLD_NF", "KLD_R"]     # hps.do_sample (:343-344)


class ResultLogger(object):
    """``ResultLogger(path, columns, append=False)``: a tab-separated log.  Format contract (``utils.py:96-107``): the
    header is written only when the file is created and carries NO trailing newline; every row is written as a newline
    followed by the tab-joined ``str.format`` of the values -- files written by the reference and by this class are
    byte-identical, so a run continued here appends cleanly to a reference log."""

    def __init__(self, path: str, columns: Iterable[str], append: bool = False):
        self.columns: List[str] = [str(c) for c in columns]
        self._fh = open(path, "a" if append else "w")
        if not append:
            self._fh.write("\t".join(self.columns))

    def log(self, run_info: Dict[str, object]) -> None:
        """One row; a column missing from ``run_info`` raises ``KeyError`` (as the reference), extra keys are ignored."""
        cells = []
        for name in self.columns:
            cells.append("{0}".format(run_info[name]))
        self._fh.write("\n" + "\t".join(cells))
        self._fh.flush()

    def close(self) -> None:
        if not self._fh.closed:
            self._fh.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
