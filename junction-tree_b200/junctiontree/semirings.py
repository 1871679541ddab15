"""Other distributive laws behind the ``SumProduct`` operator surface.

The reference anticipates them (``sum_product.py:2-3`` names the class after its law,
``junctiontree.py:300-305``: "no other distributive laws implemented currently") but ships only
sum-product.  Here every law is a semiring of the same sm_100a kernels (``JT_SR_*`` in
``include/jt_b200.h``): the product and the running reduction of the projection task are
template parameters, so collect / distribute / marginalise, uniform mode, evidence slicing and
the batch pipeline work unchanged.

================  ==============  ==========  ===============================================
law               reduction       product     propagate() returns, per factor scope
================  ==============  ==========  ===============================================
``SumProduct``    sum             ``*``       unnormalised marginals (sum = Z)
``MaxProduct``    max             ``*``       max-marginals: value of the best joint state
                                              consistent with each entry (argmax = MAP state)
``LogSumExp``     logaddexp       ``+``       log of the sum-product result, from log potentials
``MaxSum``        max             ``+``       log of the max-product result, from log potentials
================  ==============  ==========  ===============================================

Each class keeps the plugin hook of the reference: ``MaxProduct(my_einsum)`` routes every
operator through ``my_einsum`` (which must implement that law) instead of the device.
``MaxProduct`` requires non-negative potentials (max and product only form a semiring there).
"""

from . import _native
from .sum_product import SumProduct


class MaxProduct(SumProduct):
    ''' Max-product distributive law (max-marginals / MAP) '''
    semiring_flag = _native.JT_SR_MAX_PRODUCT
    name = "max-product"


class LogSumExp(SumProduct):
    ''' Sum-product in the log domain: potentials are log values, sums are logaddexp '''
    semiring_flag = _native.JT_SR_LOG_SUM_EXP
    name = "log-sum-exp"

    def ratio(self, new, old):
        raise NotImplementedError("the Hugin ratio is defined for product-domain laws only")


class MaxSum(SumProduct):
    ''' Max-product in the log domain '''
    semiring_flag = _native.JT_SR_MAX_SUM
    name = "max-sum"

    def ratio(self, new, old):
        raise NotImplementedError("the Hugin ratio is defined for product-domain laws only")


#: module-level instances bound to the device kernels, like ``computation.sum_product``
max_product = MaxProduct()
log_sum_exp = LogSumExp()
max_sum = MaxSum()
