"""Host-side mirror of the reference's model-wrapper API (reference: models/*.py).

`import models.llava` etc. at the repo root re-export these modules so that the reference's
drivers (chair_test/chair_test.py:9-19, pope_test/pope_test.py:18-22) work unchanged.
"""
