"""Minimal stand-in for the `speechbrain` package, used ONLY by
oracle/gen_golden.py so that /root/reference/src/modules/*.py and
/root/reference/src/utils/data_utils.py import unmodified in this container
(SpeechBrain is not installed and there is no network).  Not product code."""
