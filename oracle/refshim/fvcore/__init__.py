"""Stand-in for `fvcore` (absent from this image): the reference imports
`fvcore.nn` at module import time (coolchic/component/core/coolchic.py:16) but only
uses it in the *encoder's* MAC counter.  TEST INFRASTRUCTURE ONLY."""
