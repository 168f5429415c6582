"""Run-time knobs of Dropout Decoding; same names, defaults and mutability as the reference's
models/config.py:1-4.  The drivers mutate this dict (chair_test/chair_test.py:163-175); it is read
at call time.

voting_numbers : mask rates of the K ensemble members (K = len)
use_avg        : aggregate by logit mean (select_by_average) instead of max-vote
use_random     : [True] -> "epis_no_overlap" masks (no relevant-token restore), LLaVA-NeXT only
"""
settings = {}
settings['voting_numbers'] = [0.3, 0.5, 0.7]
settings['use_avg'] = False
settings['use_random'] = [False]
