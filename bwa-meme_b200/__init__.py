"""bwa-meme_b200: B200-native P-RMI seeding + banded-SW extension (the `bwa-meme mem -7` hot path).

The directory name carries a hyphen (it is the project name), so import it through
`__graft_entry__.load_pkg()` which registers it as module `bwa_meme_b200`.
"""
from . import synth, index_build, prmi_train  # noqa: F401
