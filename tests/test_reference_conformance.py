"""Build-container-only conformance tests against the REAL reference (skipped where /root/reference is absent, e.g.
on the GPU box): same seed => identical parameter names, shapes and initial values; patch_torecsys() makes the
reference's own model classes build from the drop-ins; convert() swaps modules in place sharing parameters."""
import pytest
import torch
import torch.nn as nn

from oracle.ref_shim import load_reference, reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason='reference tree not present')


@pytest.fixture(scope='module')
def ref():
    return load_reference()


@pytest.fixture()
def trs():
    import torecsys_b200 as t
    yield t
    t.unpatch_torecsys()


def _same_state(a: nn.Module, b: nn.Module):
    sa, sb = a.state_dict(), b.state_dict()
    assert list(sa.keys()) == list(sb.keys())
    for k in sa:
        assert sa[k].shape == sb[k].shape, k
        assert torch.equal(sa[k].rename(None), sb[k].rename(None)), k


CASES = [
    ('FMLayer', (0.1,), {}),
    ('FFMLayer', (6,), {'dropout_p': 0.2}),
    ('CrossNetworkLayer', (16, 3), {}),
    ('CINLayer', (8, 6, 3, [8, 4]), {}),
    ('CINLayer', (8, 6, 1, [8, 4]), {'is_direct': True, 'use_batchnorm': False}),
    ('InnerProductNetworkLayer', (7,), {}),
    ('BilinearInteractionLayer', (8, 5), {'bilinear_type': 'all'}),
    ('BilinearInteractionLayer', (8, 5), {'bilinear_type': 'each'}),
    ('AFMLayer', (8, 5, 4), {'dropout_p': 0.3}),
    ('DNNLayer', (12, 2, [16, 8]), {'dropout_p': [0.1, 0.2]}),
    ('OuterProductNetworkLayer', (8, 5), {'kernel_type': 'mat'}),
    ('OuterProductNetworkLayer', (8, 5), {'kernel_type': 'vec'}),
    ('OuterProductNetworkLayer', (8, 5), {'kernel_type': 'num'}),
    ('SENETLayer', (6, 2), {'squared': False}),
    ('CENLayer', (4, 3), {}),
]


@pytest.mark.parametrize('name,args,kwargs', CASES)
def test_layers_initialise_identically(ref, trs, name, args, kwargs):
    torch.manual_seed(7)
    theirs = getattr(ref.layers, name)(*args, **kwargs)
    torch.manual_seed(7)
    ours = getattr(trs, name)(*args, **kwargs)
    _same_state(theirs, ours)
    assert type(ours).__name__ == type(theirs).__name__
    if hasattr(theirs, 'inputs_size'):
        assert ours.inputs_size == theirs.inputs_size and ours.outputs_size == theirs.outputs_size


def test_inputs_initialise_identically(ref, trs):
    fs = [16, 32, 48, 1600]
    for cls, args in (('MultiIndicesEmbedding', (8, fs)), ('MultiIndicesFieldAwareEmbedding', (4, fs)),
                      ('SingleIndexEmbedding', (8, 100))):
        torch.manual_seed(3)
        theirs = getattr(ref.inputs.base, cls)(*args)
        torch.manual_seed(3)
        ours = getattr(trs, cls)(*args)
        _same_state(theirs, ours)
        assert len(ours) == len(theirs)
        if hasattr(theirs, 'offsets'):
            assert ours.offsets.names == theirs.offsets.names
            assert torch.equal(ours.offsets.rename(None), theirs.offsets.rename(None))


MODELS = [
    ('FactorizationMachineModel', (), {'use_bias': True, 'dropout_p': 0.1}),
    ('DeepFactorizationMachineModel', (8, 6, [16, 16, 16]), {'fm_dropout_p': 0.1}),
    ('DeepAndCrossNetworkModel', (8, 6, 4, [32, 16, 8], 3), {}),
    ('XDeepFactorizationMachineModel', (8, 6, [8, 4], [16, 16]), {}),
    ('FieldAwareFactorizationMachineModel', (6,), {'dropout_p': 0.1}),
    ('ProductNeuralNetworkModel', (8, 6, [16, 16]), {'prod_method': 'inner'}),
    ('ProductNeuralNetworkModel', (8, 6, [16, 16]), {'prod_method': 'outer', 'kernel_type': 'vec'}),
    ('FeatureImportanceAndBilinearFeatureInteractionNetwork', (8, 6, 2, 1, [16, 16]), {'bilinear_type': 'each'}),
    ('AttentionalFactorizationMachineModel', (8, 6, 4), {'dropout_p': 0.1}),
    ('NeuralFactorizationMachineModel', (8, [16, 16]), {'fm_dropout_p': 0.1}),
    ('FactorizationMachineSupportedNeuralNetworkModel', (8, 6, 1, [16, 16]), {}),
    ('DeepFieldAwareFactorizationMachineModel', (8, 6, 2, [16, 16]), {'ffm_dropout_p': 0.1}),
    ('FieldAttentiveDeepFieldAwareFactorizationMachineModel', (8, 4, 1, [16, 16], 3), {}),
]


@pytest.mark.parametrize('name,args,kwargs', MODELS)
def test_models_initialise_identically(ref, trs, name, args, kwargs):
    torch.manual_seed(11)
    theirs = getattr(ref.models, name)(*args, **kwargs)
    torch.manual_seed(11)
    ours = getattr(trs, name)(*args, **kwargs)
    _same_state(theirs, ours)


@pytest.mark.parametrize('name,args,kwargs', MODELS)
def test_patched_reference_models_build_from_drop_ins(ref, trs, name, args, kwargs):
    """models/ctr/*.py bind the layer classes at import time; after patch_torecsys() the UNCHANGED reference model
    code constructs B200 drop-in layers, with the same parameters as before."""
    torch.manual_seed(5)
    before = getattr(ref.models, name)(*args, **kwargs)
    assert trs.patch_torecsys() > 0
    torch.manual_seed(5)
    after = getattr(ref.models, name)(*args, **kwargs)
    _same_state(before, after)
    kinds = {type(m).__module__.split('.')[0] for m in after.modules() if type(m).__name__.endswith('Layer')}
    assert kinds == {'torecsys_b200'}, kinds
    assert trs.unpatch_torecsys() > 0
    again = getattr(ref.models, name)(*args, **kwargs)
    assert all(type(m).__module__.startswith('torecsys.') for m in again.modules()
               if type(m).__name__.endswith('Layer'))


def test_convert_swaps_modules_in_place_and_shares_parameters(ref, trs):
    fs = [16, 32, 48]
    feat = ref.inputs.base.MultiIndicesEmbedding(1, fs)
    emb = ref.inputs.base.MultiIndicesEmbedding(8, fs)
    feat.set_schema(['idx'])
    emb.set_schema(['idx'])
    model = ref.models.XDeepFactorizationMachineModel(8, 3, [8, 4], [16, 16])
    seq = ref.Sequential(ref.inputs.Inputs({'feat_inputs': feat, 'emb_inputs': emb}), model)
    before = {k: v for k, v in seq.named_parameters()}
    conv = trs.convert(seq)
    after = {k: v for k, v in conv.named_parameters()}
    assert list(before) == list(after)
    for k in before:
        assert before[k] is after[k], k                        # the very same Parameter objects
    assert type(conv._model.cin).__module__ == 'torecsys_b200.layers'
    assert type(conv._model.deep).__module__ == 'torecsys_b200.layers'
    assert type(conv._inputs.schema['emb_inputs']).__module__ == 'torecsys_b200.inputs'
    assert conv._inputs.schema['emb_inputs'] is conv._inputs._modules['emb_inputs']
    assert list(conv.state_dict().keys()) == list(seq.state_dict().keys())
