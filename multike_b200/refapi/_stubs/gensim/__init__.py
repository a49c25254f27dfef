"""Stand-in for gensim when it is not installed: the reference's utils.py imports Word2Vec at module
level but only uses it for character-level vectors of out-of-vocabulary words."""
