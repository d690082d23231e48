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
    """``ResultLogger(path, columns, append=False)``: a tab-separated log whose header line is written WITHOUT a trailing
    newline (only when the file is created) and whose rows each start with one (``utils.py:96-107``) -- files written by
    the reference and by this class are byte-identical, so a run continued here appends cleanly to a reference log."""

    def __init__(self, path: str, columns: Iterable[str], append: bool = False):
        self.columns: List[str] = list(columns)
        mode = "a" if append else "w"
        self.f_log = open(path, mode)
        if mode == "w":
            self.f_log.write("\t".join(self.columns))

    def __del__(self):
        try:
            self.f_log.close()
        except Exception:
            pass

    def log(self, run_info: Dict[str, object]) -> None:
        run_strings = ["{0}".format(run_info[lc]) for lc in self.columns]      # KeyError for a missing column, as the reference
        self.f_log.write("\n")
        self.f_log.write("\t".join(run_strings))
        self.f_log.flush()
