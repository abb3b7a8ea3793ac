"""Hit compaction + LArPix packet builder (reference fee.py:84-359 export_to_hdf5).

CPU: the Python restatement (oracle/packets_oracle.py) reproduces what the reference's own export_to_hdf5
produced on the committed fixtures.  GPU: the CUDA builder (through the C ABI / the fee.export_to_hdf5 drop-in)
is identical to the restatement, field for field, and to the reference fixtures."""
import numpy as np
import pytest

import packets_util as pu
from packets_util import po


@pytest.mark.parametrize("name", pu.CASES)
def test_oracle_reproduces_reference_packets(name):
    z = pu.load(name)
    packets, ds = po.export_packets(pu.tables_from_npz(z), bad_channels=pu.bad_channels_from_npz(z), **pu.inputs(z))
    assert (packets["packet_type"] == po.PT_DATA).sum() > 500
    pu.check_against_reference(packets, ds, z)


def test_oracle_rollover_case_has_rollovers():
    z = pu.load("module0_rollover_bad")
    t = pu.tables_from_npz(z)
    packets, _ = po.export_packets(t, bad_channels=pu.bad_channels_from_npz(z), **pu.inputs(z))
    # event start times beyond the 31-bit clock period: timestamps wrapped, sync + trigger packets present
    assert (z["in_event_start_times"] / t["clock_cycle"] > t["clock_reset_period"]).any()
    assert (packets["timestamp"][packets["packet_type"] != po.PT_TIMESTAMP] < t["clock_reset_period"]).all()
    assert (packets["packet_type"] == po.PT_SYNC).any() and (packets["packet_type"] == po.PT_TRIGGER).any()


def test_parity_is_odd_over_the_word():
    for chip, ch, ts, dw in ((11, 3, 12345, 87), (255, 63, 2**31 - 1, 255), (0, 0, 0, 0)):
        word = (chip << 2) | (ch << 10) | (ts << 16) | (1 << 47) | (dw << 48)
        p = po.data_parity(chip, ch, ts, 1, dw)
        assert (bin(word).count("1") + p) % 2 == 1
