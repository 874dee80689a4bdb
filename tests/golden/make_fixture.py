"""Packs the reference's golden-vector INPUT (src/test/java/SevenZip/firefox.exe, the file
LzmaAloneTest.java:27-38 compresses with 12 switch sets) into tests/golden/firefox.exe.xz so
that the 12 (length, md5) vectors can be checked where /root/reference does not exist (the
GPU box).  Test data, not reference source; stored xz-compressed (Python's lzma module) to
keep the history small.  The md5 of the unpacked bytes is asserted by the tests."""
import hashlib
import lzma
import os

SRC = "/root/reference/src/test/java/SevenZip/firefox.exe"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "firefox.exe.xz")

if __name__ == "__main__":
    data = open(SRC, "rb").read()
    assert hashlib.md5(data).hexdigest() == "5744fff8e72d105c138dae9e17bb29fe"
    with open(DST, "wb") as f:
        f.write(lzma.compress(data, preset=9 | lzma.PRESET_EXTREME))
    assert lzma.decompress(open(DST, "rb").read()) == data
    print(DST, os.path.getsize(DST))
