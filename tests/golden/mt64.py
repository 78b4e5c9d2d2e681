"""std::mt19937_64 in pure Python (the golden 1 Mbp vector of SURVEY.md §8c is
defined with `std::mt19937_64 rng(1234)`; numpy only ships the 32-bit variant)."""


class mt19937_64:
    NN, MM = 312, 156
    MATRIX_A, UM, LM = 0xB5026F5AA96619E9, 0xFFFFFFFF80000000, 0x7FFFFFFF
    M64 = 0xFFFFFFFFFFFFFFFF

    def __init__(self, seed):
        mt = [0] * self.NN
        mt[0] = seed & self.M64
        for i in range(1, self.NN):
            mt[i] = (6364136223846793005 * (mt[i - 1] ^ (mt[i - 1] >> 62)) + i) & self.M64
        self.mt, self.mti = mt, self.NN

    def next(self):
        mt = self.mt
        if self.mti >= self.NN:
            for i in range(self.NN):
                x = (mt[i] & self.UM) | (mt[(i + 1) % self.NN] & self.LM)
                mt[i] = mt[(i + self.MM) % self.NN] ^ (x >> 1) ^ (self.MATRIX_A if x & 1 else 0)
            self.mti = 0
        x = mt[self.mti]
        self.mti += 1
        x ^= (x >> 29) & 0x5555555555555555
        x ^= (x << 17) & 0x71D67FFFEDA60000
        x ^= (x << 37) & 0xFFF7EEE000000000
        x ^= x >> 43
        return x
