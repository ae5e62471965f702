"""Running average with a reference value.

Same semantics as the reference's ``Averager``
(/root/reference/cemc/mcmc/averager.py:2-66): values are accumulated as
``value / ref_value``; ``mean`` multiplies back.  On the GPU path the sums are
produced by the kernels (accumulator slots CEMC_ACC_E / CEMC_ACC_E2) and
loaded with :meth:`set_sums`.
"""


class Averager(object):
    def __init__(self, ref_value=1.0):
        self._ref_value = float(ref_value)
        self._n_samples = 0.0
        self._mean = 0.0

    def __iadd__(self, value):
        if isinstance(value, Averager):
            self._mean += value._mean * (value._ref_value / self._ref_value)
            self._n_samples += value._n_samples
            return self
        self._n_samples += 1.0
        self._mean += value / self._ref_value
        return self

    def __add__(self, other):
        ratio = (other._ref_value / self._ref_value)
        new_obj = Averager(ref_value=self._ref_value)
        new_obj._mean = self._mean + other._mean * ratio
        new_obj._n_samples = self._n_samples + other._n_samples
        return new_obj

    def __itruediv__(self, number):
        return self.__idiv__(number)

    def __idiv__(self, number):
        self._mean /= float(number)
        self._n_samples /= float(number)
        return self

    def clear(self):
        self._n_samples = 0
        self._mean = 0.0

    def set_sums(self, scaled_sum, n_samples):
        """Load sums accumulated on the device (sum of value/ref, count)."""
        self._mean = float(scaled_sum)
        self._n_samples = float(n_samples)

    @property
    def ref_value(self):
        return self._ref_value

    @property
    def mean(self):
        if self._n_samples == 0:
            return self._mean * self._ref_value
        return (self._mean / self._n_samples) * self._ref_value
