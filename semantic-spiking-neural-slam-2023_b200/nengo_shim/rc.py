"""``nengo.rc`` stand-in: the drivers only assign ``rc['progress']['progress_bar']``
(``experiments/run_slam.py:197``)."""
import collections

rc = collections.defaultdict(dict)
rc["precision"]["bits"] = 32
rc["progress"]["progress_bar"] = "none"
