"""Encoder factory over the reference's argument namespace (contract: reference encoders/build.py:7-152).

Table-driven: each encoder family lists which constructor keyword reads which `args` attribute; entries that are
callables compute the value from `args`.  `tests/test_cli_contract.py::test_build_encoder_factory_matches_reference`
pins class, parameter names/shapes, caller-visible properties and the seeded initial weights to the reference's factory."""

_LEGACY_NAMES = (('transformer_enc_d_model', 'transformer_d_model'), ('transformer_dec_d_model', 'transformer_d_model'),
                 ('transformer_enc_d_ff', 'transformer_d_ff'), ('transformer_enc_n_heads', 'transformer_n_heads'))


def _input_dim(a):
    return a.input_dim if a.input_type == 'speech' else a.emb_dim


def _last_proj_dim(a):
    return a.transformer_dec_d_model if 'transformer' in a.dec_type else 0


_SHARED = dict(input_dim=_input_dim, enc_type='enc_type', last_proj_dim=_last_proj_dim, n_layers='enc_n_layers',
               n_layers_sub1='enc_n_layers_sub1', n_layers_sub2='enc_n_layers_sub2', dropout_in='dropout_in',
               dropout='dropout_enc', subsample='subsample', subsample_type='subsample_type', n_stacks='n_stacks',
               n_splices='n_splices', task_specific_layer='task_specific_layer')

_FORMER = dict(_SHARED, n_heads='transformer_enc_n_heads', d_model='transformer_enc_d_model', d_ff='transformer_enc_d_ff',
               ffn_bottleneck_dim='transformer_ffn_bottleneck_dim', pe_type='transformer_enc_pe_type',
               layer_norm_eps='transformer_layer_norm_eps', dropout_att='dropout_att', dropout_layer='dropout_enc_layer',
               param_init='transformer_param_init', clamp_len='transformer_enc_clamp_len',
               lookahead='transformer_enc_lookaheads', chunk_size_left='lc_chunk_size_left',
               chunk_size_current='lc_chunk_size_current', chunk_size_right='lc_chunk_size_right', streaming_type='lc_type')

_TRANSFORMER = dict(_FORMER, ffn_activation='transformer_ffn_activation')
_CONFORMER = dict(_FORMER, ffn_activation=lambda a: 'swish', kernel_size='conformer_kernel_size',
                  normalization='conformer_normalization')
_RNN = dict(_SHARED, n_units='enc_n_units', n_projs='enc_n_projs', bidir_sum_fwd_bwd='bidirectional_sum_fwd_bwd',
            param_init='param_init', chunk_size_current='lc_chunk_size_left',      # sic: the reference passes _left here
            chunk_size_right='lc_chunk_size_right', cnn_lookahead='cnn_lookahead', rsp_prob='rsp_prob_enc')

_CONV = dict(in_channel='conv_in_channel', channels='conv_channels', kernel_sizes='conv_kernel_sizes',
             strides='conv_strides', poolings='conv_poolings', normalization='conv_normalization', param_init='param_init',
             dropout=lambda a: 0., residual=lambda a: False,
             bottleneck_dim=lambda a: a.transformer_enc_d_model if 'former' in a.enc_type else a.conv_bottleneck_dim)


def _kwargs(table, args):
    return {k: (src(args) if callable(src) else getattr(args, src)) for k, src in table.items()}


def build_encoder(args):
    for new, old in _LEGACY_NAMES:          # checkpoints trained with the old option names (reference build.py:25-32)
        if not hasattr(args, new) and hasattr(args, old):
            setattr(args, new, getattr(args, old))
    if args.enc_type in ('tds', 'gated_conv'):
        raise NotImplementedError("enc_type=%r is outside the B200 hot path (SURVEY.md section 2)" % args.enc_type)
    conv = None
    if 'conv' in args.enc_type:
        from .conv import ConvEncoder
        assert args.n_stacks == 1 and args.n_splices == 1
        conv = ConvEncoder(args.input_dim, **_kwargs(_CONV, args))
    if 'transformer' in args.enc_type:
        from .transformer import TransformerEncoder
        return TransformerEncoder(frontend_conv=conv, **_kwargs(_TRANSFORMER, args))
    if 'conformer' in args.enc_type:
        from .conformer import ConformerEncoder
        return ConformerEncoder(frontend_conv=conv, **_kwargs(_CONFORMER, args))
    from .rnn import RNNEncoder               # LSTM / BLSTM families and the CNN-only encoder (enc_type == 'conv')
    return RNNEncoder(frontend_conv=conv, **_kwargs(_RNN, args))
