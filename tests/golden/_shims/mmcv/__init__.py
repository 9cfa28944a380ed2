"""Import shim used ONLY by tests/golden/make_golden.py: the reference's CAVP modules import mmcv
(not installed, not vendored) for *wiring* only -- ConvModule is conv -> bn -> activation with
sub-module names `conv`, `bn`, `activate`; all arithmetic is torch.nn."""
