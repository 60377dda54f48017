"""Import alias: the implementation lives in `graphless-neural-networks_b200/` (the directory name
the project layout prescribes; a hyphen is not importable), so `import glnn_b200.<module>` resolves
there."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                          "graphless-neural-networks_b200")]
__version__ = "0.1.0"
