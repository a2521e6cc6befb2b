"""Record layouts on the drop-in boundary (unchanged from the reference).

* `instruction_dtype`      -- wfsim/strax_interface.py:25-42 (70-byte packed rows)
* `truth_extra_dtype`      -- wfsim/strax_interface.py:49-73
* `extra_truth_dtype_per_pmt` -- wfsim/strax_interface.py:77-116
* `raw_record_dtype`       -- strax.raw_record_dtype (third party; 244-byte layout)

The C-ABI library reads/writes these exact packed layouts (see include/wfsim_b200.h).
"""
import numpy as np

RECORD_LENGTH = 110

instruction_dtype = [
    (('Waveform simulator event number.', 'event_number'), np.int32),
    (('Quanta type (S1 photons or S2 electrons)', 'type'), np.int8),
    (('Time of the interaction [ns]', 'time'), np.int64),
    (('X position of the cluster [cm]', 'x'), np.float32),
    (('Y position of the cluster [cm]', 'y'), np.float32),
    (('Z position of the cluster [cm]', 'z'), np.float32),
    (('Number of quanta', 'amp'), np.int32),
    (('Recoil type of interaction.', 'recoil'), np.int8),
    (('Energy deposit of interaction', 'e_dep'), np.float32),
    (('Total energy deposit in the sensitive volume', 'tot_e'), np.float32),
    (('Eventid like in geant4 output rootfile', 'g4id'), np.int32),
    (('Volume id giving the detector subvolume', 'vol_id'), np.int32),
    (('Local field [ V / cm ]', 'local_field'), np.float64),
    (('Number of excitons', 'n_excitons'), np.int32),
    (('X position of the primary particle [cm]', 'x_pri'), np.float32),
    (('Y position of the primary particle [cm]', 'y_pri'), np.float32),
    (('Z position of the primary particle [cm]', 'z_pri'), np.float32),
]

_TRUTH_COUNTERS_INT = ['n_photon', 'n_pe', 'n_photon_trigger', 'n_pe_trigger']
_TRUTH_COUNTERS_FLT = ['raw_area', 'raw_area_trigger']

_TRUTH_TAIL = [
    (('Arrival time of the first photon [ns]', 't_first_photon'), np.float64),
    (('Arrival time of the last photon [ns]', 't_last_photon'), np.float64),
    (('Mean time of the photons [ns]', 't_mean_photon'), np.float64),
    (('Standard deviation of photon arrival times [ns]', 't_sigma_photon'), np.float64),
    (('X field-distorted mean position of the electrons [cm]', 'x_mean_electron'), np.float32),
    (('Y field-distorted mean position of the electrons [cm]', 'y_mean_electron'), np.float32),
    (('Arrival time of the first electron [ns]', 't_first_electron'), np.float64),
    (('Arrival time of the last electron [ns]', 't_last_electron'), np.float64),
    (('Mean time of the electrons [ns]', 't_mean_electron'), np.float64),
    (('Standard deviation of electron arrival times [ns]', 't_sigma_electron'), np.float64),
]

_DESCR = {
    'n_photon': 'Number of photons reaching PMT',
    'n_pe': 'Number of photons + dpe passing',
    'n_photon_trigger': 'Number of photons passing trigger',
    'n_pe_trigger': 'Number of photons + dpe passing trigger',
    'raw_area': 'Raw area in pe',
    'raw_area_trigger': 'Raw area in pe passing trigger',
}

truth_extra_dtype = (
    [(('End time of the interaction [ns]', 'endtime'), np.int64),
     (('Number of simulated electrons', 'n_electron'), np.int32)]
    + [((_DESCR[f], f), np.int32) for f in _TRUTH_COUNTERS_INT]
    + [((_DESCR[f], f), np.float64) for f in _TRUTH_COUNTERS_FLT]
    + [((_DESCR[f] + ' (bottom)', f + '_bottom'), np.int32) for f in _TRUTH_COUNTERS_INT]
    + [((_DESCR[f] + ' (bottom)', f + '_bottom'), np.float64) for f in _TRUTH_COUNTERS_FLT]
    + _TRUTH_TAIL)


# strax_interface.py:44-45: optical (externally supplied photons) instructions carry the index range of
# their photons in the channel / timing lists
optical_extra_dtype = [(('first optical input index', '_first'), np.int32),
                       (('last optical input index +1', '_last'), np.int32)]


def extra_truth_dtype_per_pmt(n_pmt):
    """Truth layout; total/bottom split when `n_pmt` is falsy, per-PMT arrays otherwise."""
    if not n_pmt:
        return truth_extra_dtype
    return (
        [(('End time of the interaction [ns]', 'endtime'), np.int64),
         (('Number of simulated electrons', 'n_electron'), np.int32)]
        + [((_DESCR[f], f + '_per_pmt'), (np.int32, n_pmt)) for f in _TRUTH_COUNTERS_INT]
        + [((_DESCR[f], f + '_per_pmt'), (np.float64, n_pmt)) for f in _TRUTH_COUNTERS_FLT]
        + [((_DESCR[f] + ' (total)', f), np.int32) for f in _TRUTH_COUNTERS_INT]
        + [((_DESCR[f] + ' (total)', f), np.float64) for f in _TRUTH_COUNTERS_FLT]
        + _TRUTH_TAIL)


def raw_record_dtype(samples_per_record=RECORD_LENGTH):
    return np.dtype([
        (('Start time since unix epoch [ns]', 'time'), np.int64),
        (('Length of the interval in samples', 'length'), np.int32),
        (('Width of one sample [ns]', 'dt'), np.int16),
        (('Channel/PMT number', 'channel'), np.int16),
        (('Length of pulse to which the record belongs (without zero-padding)', 'pulse_length'), np.int32),
        (('Fragment number in the pulse', 'record_i'), np.int16),
        (('Baseline determined by the digitizer (if this is supported)', 'baseline'), np.int16),
        (('Waveform data in raw ADC counts', 'data'), np.int16, samples_per_record)])


def truth_dtype(per_pmt_n=False):
    return np.dtype(instruction_dtype + extra_truth_dtype_per_pmt(per_pmt_n))


assert np.dtype(instruction_dtype).itemsize == 70
assert truth_dtype().itemsize == 218
assert raw_record_dtype().itemsize == 244
