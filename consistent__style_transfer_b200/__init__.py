"""B200-native Word Mover's Distance engine: drop-in for the content-preservation scoring path
of iptmt/consistent__style_transfer (src/wmd.py, evaluate/auto/content_preserve.py)."""
