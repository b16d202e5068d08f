"""Command-line contract of the encoders: ``add_args(parser, args)`` and ``define_name(dir_name, args)``.

The reference's trainer discovers encoder options and the experiment-directory name through these two static methods
(neural_sp/bin/args_asr.py:87-100; encoders/conformer.py:113-191, transformer.py:275-345, conv.py:103-135,
rnn.py:214-257).  They are part of the drop-in boundary: option names, types, defaults, choices and the
directory-name grammar must match, otherwise existing recipes / checkpoints directories do not resolve.

Implemented as option tables + one naming grammar instead of per-class argparse code; `tests/test_cli_contract.py`
pins both against tables extracted from the unmodified reference (tests/golden/cli_contract.json).
"""
from distutils.util import strtobool

_ACTS = ['relu', 'gelu', 'gelu_accurate', 'glu', 'swish']

# (flag, type, default, choices-or-None)
CONV_OPTS = [
    ('--conv_in_channel', int, 1, None),
    ('--conv_channels', str, "", None),
    ('--conv_kernel_sizes', str, "", None),
    ('--conv_strides', str, "", None),
    ('--conv_poolings', str, "", None),
    ('--conv_normalization', str, '', ['', 'layer_norm', 'batch_norm']),
    ('--conv_bottleneck_dim', int, 0, None),
]


def _former_common(default_act, with_input_bottleneck):
    opts = [('--transformer_ffn_bottleneck_dim', int, 0, None)]
    if with_input_bottleneck:
        opts.append(('--transformer_input_bottleneck_dim', int, 0, None))
    opts += [('--transformer_layer_norm_eps', float, 1e-12, None),
             ('--transformer_ffn_activation', str, default_act, _ACTS),
             ('--transformer_param_init', str, 'xavier_uniform', ['xavier_uniform', 'pytorch'])]
    return opts


_STREAMING_FORMER = [
    ('--transformer_enc_lookaheads', str, "0_0_0_0_0_0_0_0_0_0_0_0", None),
    ('--lc_chunk_size_left', str, "0", None),
    ('--lc_chunk_size_current', str, "0", None),
    ('--lc_chunk_size_right', str, "0", None),
    ('--lc_type', str, 'reshape', ['reshape', 'mask']),
]

TRANSFORMER_OPTS = [
    ('--transformer_enc_d_model', int, 256, None),
    ('--transformer_enc_d_ff', int, 2048, None),
    ('--transformer_enc_n_heads', int, 4, None),
    ('--transformer_enc_pe_type', str, 'add', ['add', 'none', 'relative', 'relative_xl']),
    ('--dropout_enc_layer', float, 0.0, None),
    ('--transformer_enc_clamp_len', int, -1, None),
] + _STREAMING_FORMER

CONFORMER_OPTS = [
    ('--transformer_enc_d_model', int, 256, None),
    ('--transformer_enc_d_ff', int, 2048, None),
    ('--transformer_enc_n_heads', int, 4, None),
    ('--transformer_enc_pe_type', str, 'relative', ['relative', 'relative_xl', 'none']),
    ('--conformer_kernel_size', int, 31, None),
    ('--conformer_normalization', str, 'batch_norm', ['batch_norm', 'group_norm', 'layer_norm']),
    ('--dropout_enc_layer', float, 0.0, None),
    ('--transformer_enc_clamp_len', int, -1, None),
] + _STREAMING_FORMER

RNN_OPTS = [
    ('--enc_n_units', int, 512, None),
    ('--enc_n_projs', int, 0, None),
    ('--bidirectional_sum_fwd_bwd', strtobool, False, None),
    ('--lc_chunk_size_left', str, "-1", None),
    ('--lc_chunk_size_right', str, "0", None),
    ('--cnn_lookahead', strtobool, True, None),
    ('--rsp_prob_enc', float, 0.0, None),
]


def _add(group, opts):
    for flag, typ, default, choices in opts:
        kw = dict(type=typ, default=default)
        if choices is not None:
            kw['choices'] = choices
        group.add_argument(flag, **kw)


def conv_add_args(parser, args):
    _add(parser.add_argument_group("CNN encoder"), CONV_OPTS)
    return parser


def _former_add_args(parser, args, title, own, default_act, with_input_bottleneck):
    group = parser.add_argument_group(title)
    if 'conv' in args.enc_type:
        parser = conv_add_args(parser, args)
    if not hasattr(args, 'transformer_layer_norm_eps'):     # shared with the Transformer decoder's options
        _add(group, _former_common(default_act, with_input_bottleneck))
    _add(group, own)
    return parser


def transformer_add_args(parser, args):
    return _former_add_args(parser, args, "Transformer encoder", TRANSFORMER_OPTS, 'relu', True)


def conformer_add_args(parser, args):
    return _former_add_args(parser, args, "Transformer encoder", CONFORMER_OPTS, 'swish', False)


def rnn_add_args(parser, args):
    group = parser.add_argument_group("RNN encoder")
    parser = conv_add_args(parser, args)
    _add(group, RNN_OPTS)
    return parser


# ---------------------------------------------------------------------------------------------
# experiment-directory names
# ---------------------------------------------------------------------------------------------
def _last(v):
    return int(str(v).split('_')[-1])


def conv_define_name(dir_name, args):
    assert 'conv' in args.enc_type
    base = args.enc_type.replace('conv_', '')
    n_blocks = len(args.conv_channels.split('_')) if args.conv_channels else 0
    if n_blocks == 0:
        return base
    return 'conv%dL%s%s' % (n_blocks, args.conv_normalization or '', base)


def _former_define_name(dir_name, args, middle):
    if 'conv' in args.enc_type:
        dir_name = conv_define_name(dir_name, args)
    parts = ['%ddmodel' % args.transformer_enc_d_model, '%ddff' % args.transformer_enc_d_ff]
    if args.transformer_ffn_bottleneck_dim > 0:
        parts.append('%dbn' % args.transformer_ffn_bottleneck_dim)
    parts += ['%dL' % args.enc_n_layers, '%dH' % args.transformer_enc_n_heads, middle]
    if args.transformer_enc_clamp_len > 0:
        parts.append('_clamp%d' % args.transformer_enc_clamp_len)
    if args.dropout_enc_layer > 0:
        parts.append('_LD' + str(args.dropout_enc_layer))
    chunks = (args.lc_chunk_size_left, args.lc_chunk_size_current, args.lc_chunk_size_right)
    lookahead = sum(int(v) for v in args.transformer_enc_lookaheads.split('_'))
    if any(_last(c) > 0 for c in chunks):
        parts.append('_chunkL%sC%sR%s_%s' % (chunks[0], chunks[1], chunks[2], args.lc_type))
    elif lookahead > 0:
        parts.append('_LA%d' % lookahead)
    return dir_name + ''.join(parts)


def transformer_define_name(dir_name, args):
    return _former_define_name(dir_name, args, 'pe' + str(args.transformer_enc_pe_type))


def conformer_define_name(dir_name, args):
    return _former_define_name(dir_name, args, 'kernel%d_%s' % (args.conformer_kernel_size, args.conformer_normalization))


def rnn_define_name(dir_name, args):
    if 'conv' in args.enc_type:
        dir_name = conv_define_name(dir_name, args)
    parts = ['%dH' % args.enc_n_units]
    if args.enc_n_projs > 0:
        parts.append('%dP' % args.enc_n_projs)
    parts.append('%dL' % args.enc_n_layers)
    if args.bidirectional_sum_fwd_bwd:
        parts.append('_sumfwdbwd')
    first = lambda v: int(str(v).split('_')[0])     # noqa: E731  (the RNN grammar looks at the FIRST chunk entry)
    if first(args.lc_chunk_size_left) > 0 or first(args.lc_chunk_size_right) > 0:
        parts.append('_chunkL%sR%s' % (args.lc_chunk_size_left, args.lc_chunk_size_right))
        if not args.cnn_lookahead:
            parts.append('_blockwise')
    if args.rsp_prob_enc > 0:
        parts.append('_RSP' + str(args.rsp_prob_enc))
    return dir_name + ''.join(parts)
