"""Drop-in replacement for the reference's ``vilbert`` package.

Put ``youtube-vln_b200/`` ahead of the YouTube-VLN checkout on ``sys.path``: ``vilbert.vilbert`` then resolves to
the B200-native implementation in this directory, while sub-modules this package does not provide
(``vilbert.optimization``, ``vilbert.vilbert_init``, ``vilbert.file_utils``) keep resolving to the reference's own
files through the extended package path.
"""
import pkgutil

__path__ = pkgutil.extend_path(__path__, __name__)
