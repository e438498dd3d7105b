"""``Read`` value class with the reference's surface (scripts/read.py:5-27).

Nothing on the recruitment path imports it (reads enter as NCRF records); it is kept so
that code written against the reference's module keeps working."""


class Read:
    def __init__(self, id, seq=None, simulated=False):
        self.id, self.seq = id, seq
        if simulated:  # SimLoRD read-name fields (scripts/read.py:8-16)
            f = self.id.split("_")
            self.numb = int(f[1])
            self.length = int(f[2].split("=")[1][:-2])
            self.start_pos = int(f[3].split("=")[1])
            self.n_errors = int(f[6].split("=")[1])
            self.error_rate = float(f[9].split("=")[1])
            self.mult = float(f[-1].split("=")[1])

    @classmethod
    def FromBiopyRead(cls, biopy_read, simulated=False):
        return cls(biopy_read.id, str(biopy_read.seq), simulated)

    def __len__(self):
        return len(self.seq)

    def __getitem__(self, key):
        return self.seq[key]
