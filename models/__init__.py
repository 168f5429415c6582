"""Drop-in alias of the reference's `models` package: `from models.llava import
CustomLlavaForConditionalGeneration` (chair_test/chair_test.py:9-19) resolves to the B200 wrappers."""
