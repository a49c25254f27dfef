class Word2Vec:
    def __init__(self, *args, **kwargs):
        raise ImportError("gensim is needed for character-level vectors of out-of-vocabulary words "
                          "(utils.generate_word2vec_by_character_embedding); install it or pre-compute literal_vectors.npy")
