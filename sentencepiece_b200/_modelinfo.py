"""Piece strings of a serialized ModelProto (sentencepiece_model.proto:293-332) for the Python mirror's piece
output: a minimal wire-format reader (field 1 = repeated SentencePiece {1: piece, 2: score, 3: type})."""


def _varint(b, p):
    v = s = 0
    while True:
        c = b[p]
        p += 1
        v |= (c & 0x7F) << s
        if c < 0x80:
            return v, p
        s += 7


def pieces_and_unk(data):
    pieces, unk, p, n = [], -1, 0, len(data)
    while p < n:
        key, p = _varint(data, p)
        fno, wt = key >> 3, key & 7
        if wt == 2:
            ln, p = _varint(data, p)
            if fno == 1:
                q, end, piece, typ = p, p + ln, b"", 1
                while q < end:
                    k2, q = _varint(data, q)
                    f2, w2 = k2 >> 3, k2 & 7
                    if w2 == 2:
                        l2, q = _varint(data, q)
                        if f2 == 1:
                            piece = data[q:q + l2]
                        q += l2
                    elif w2 == 0:
                        v, q = _varint(data, q)
                        if f2 == 3:
                            typ = v
                    elif w2 == 5:
                        q += 4
                    elif w2 == 1:
                        q += 8
                    else:
                        raise ValueError("unsupported wire type in SentencePiece")
                if typ == 2:  # UNKNOWN
                    unk = len(pieces)
                pieces.append(piece.decode("utf-8", errors="replace"))
            p += ln
        elif wt == 0:
            _, p = _varint(data, p)
        elif wt == 5:
            p += 4
        elif wt == 1:
            p += 8
        else:
            raise ValueError("unsupported wire type in ModelProto")
    return pieces, unk
