"""``Vocab`` with the reference's attributes (utils/data.py:1-28): four special tokens with ids 0..3, then the
labels of the labels JSON in file order.  The corpus-cleaning helpers of the reference file (SEAME text
normalisation, sox segmentation; utils/data.py:61-482) are offline data preparation, not the training path."""


class Vocab(object):
    def __init__(self):
        self.PAD_TOKEN, self.SOS_TOKEN, self.EOS_TOKEN, self.OOV_TOKEN = "<PAD>", "<SOS>", "<EOS>", "<OOV>"
        self.PAD_ID, self.SOS_ID, self.EOS_ID, self.OOV_ID = 0, 1, 2, 3
        self.special_token_list = [self.PAD_TOKEN, self.SOS_TOKEN, self.EOS_TOKEN, self.OOV_TOKEN]
        self.token2id, self.id2token = {}, []
        self.label2id, self.id2label = {}, []
        for tok in self.special_token_list:
            self.add_token(tok)
            self.add_label(tok)

    def add_token(self, token):
        if token not in self.token2id:
            self.token2id[token] = len(self.id2token)
            self.id2token.append(token)

    def add_label(self, label):
        if label not in self.label2id:
            self.label2id[label] = len(self.id2label)
            self.id2label.append(label)
