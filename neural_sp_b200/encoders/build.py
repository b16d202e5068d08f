"""Encoder factory with the reference's argument namespace (reference encoders/build.py:7-152)."""


def build_encoder(args):
    # safeguard for checkpoints trained with the old option names (reference build.py:25-32)
    if not hasattr(args, 'transformer_enc_d_model') and hasattr(args, 'transformer_d_model'):
        args.transformer_enc_d_model = args.transformer_d_model
        args.transformer_dec_d_model = args.transformer_d_model
    if not hasattr(args, 'transformer_enc_d_ff') and hasattr(args, 'transformer_d_ff'):
        args.transformer_enc_d_ff = args.transformer_d_ff
    if not hasattr(args, 'transformer_enc_n_heads') and hasattr(args, 'transformer_n_heads'):
        args.transformer_enc_n_heads = args.transformer_n_heads
    conv = None
    if 'conv' in args.enc_type:
        from .conv import ConvEncoder
        assert args.n_stacks == 1 and args.n_splices == 1
        conv = ConvEncoder(args.input_dim, in_channel=args.conv_in_channel, channels=args.conv_channels,
                           kernel_sizes=args.conv_kernel_sizes, strides=args.conv_strides,
                           poolings=args.conv_poolings, dropout=0., normalization=args.conv_normalization,
                           residual=False, bottleneck_dim=args.transformer_enc_d_model
                           if ('former' in args.enc_type) else args.conv_bottleneck_dim, param_init=args.param_init)
    if args.enc_type in ('tds', 'gated_conv'):
        raise NotImplementedError("enc_type=%r is outside the B200 hot path (SURVEY.md section 2)" % args.enc_type)
    if 'former' not in args.enc_type:
        return _build_rnn(args, conv)
    common = dict(
        input_dim=args.input_dim if args.input_type == 'speech' else args.emb_dim, enc_type=args.enc_type,
        n_heads=args.transformer_enc_n_heads, n_layers=args.enc_n_layers, n_layers_sub1=args.enc_n_layers_sub1,
        n_layers_sub2=args.enc_n_layers_sub2, d_model=args.transformer_enc_d_model, d_ff=args.transformer_enc_d_ff,
        ffn_bottleneck_dim=args.transformer_ffn_bottleneck_dim, pe_type=args.transformer_enc_pe_type,
        layer_norm_eps=args.transformer_layer_norm_eps,
        last_proj_dim=args.transformer_dec_d_model if 'transformer' in args.dec_type else 0,
        dropout_in=args.dropout_in, dropout=args.dropout_enc, dropout_att=args.dropout_att,
        dropout_layer=args.dropout_enc_layer, subsample=args.subsample, subsample_type=args.subsample_type,
        n_stacks=args.n_stacks, n_splices=args.n_splices, frontend_conv=conv,
        task_specific_layer=args.task_specific_layer, param_init=args.transformer_param_init,
        clamp_len=args.transformer_enc_clamp_len, lookahead=args.transformer_enc_lookaheads,
        chunk_size_left=args.lc_chunk_size_left, chunk_size_current=args.lc_chunk_size_current,
        chunk_size_right=args.lc_chunk_size_right, streaming_type=args.lc_type)
    if 'conformer' in args.enc_type:
        from .conformer import ConformerEncoder
        return ConformerEncoder(kernel_size=args.conformer_kernel_size, normalization=args.conformer_normalization,
                                ffn_activation='swish', **common)
    if 'transformer' in args.enc_type:
        from .transformer import TransformerEncoder
        return TransformerEncoder(ffn_activation=args.transformer_ffn_activation, **common)
    raise AssertionError("unreachable")


def _build_rnn(args, conv):
    """RNN family, including the CNN-only encoder enc_type='conv' (the reference wraps it in RNNEncoder too)."""
    from .rnn import RNNEncoder
    return RNNEncoder(input_dim=args.input_dim if args.input_type == 'speech' else args.emb_dim, enc_type=args.enc_type,
                      n_units=args.enc_n_units, n_projs=args.enc_n_projs,
                      last_proj_dim=args.transformer_dec_d_model if 'transformer' in args.dec_type else 0,
                      n_layers=args.enc_n_layers, n_layers_sub1=args.enc_n_layers_sub1,
                      n_layers_sub2=args.enc_n_layers_sub2, dropout_in=args.dropout_in, dropout=args.dropout_enc,
                      subsample=args.subsample, subsample_type=args.subsample_type, n_stacks=args.n_stacks,
                      n_splices=args.n_splices, frontend_conv=conv, bidir_sum_fwd_bwd=args.bidirectional_sum_fwd_bwd,
                      task_specific_layer=args.task_specific_layer, param_init=args.param_init,
                      chunk_size_current=args.lc_chunk_size_left, chunk_size_right=args.lc_chunk_size_right,
                      cnn_lookahead=args.cnn_lookahead, rsp_prob=args.rsp_prob_enc)
