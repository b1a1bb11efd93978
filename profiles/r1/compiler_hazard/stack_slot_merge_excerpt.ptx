	// end inline asm
	mov.b32 	%r512, 0;
	// begin inline asm
	subc.u32 %r459, %r512, %r512;
	// end inline asm
	setp.eq.s32 	%p24, %r459, 0;
	selp.b32 	%r17, %r435, %r411, %p24;
	selp.b32 	%r18, %r438, %r414, %p24;
	selp.b32 	%r19, %r441, %r417, %p24;
	selp.b32 	%r20, %r444, %r420, %p24;
	selp.b32 	%r25, %r447, %r423, %p24;
	selp.b32 	%r26, %r450, %r426, %p24;
	selp.b32 	%r23, %r453, %r429, %p24;
	selp.b32 	%r24, %r456, %r432, %p24;
	cvta.to.local.u64 	%rd1, %rd13;
	st.local.v4.u32 	[%rd1], {%r17, %r18, %r19, %r20};
	st.local.v4.u32 	[%rd1+16], {%r25, %r26, %r23, %r24};
	{ // callseq 10, 0
	.param .b64 param0;
	st.param.b64 	[param0], %rd13;
	.param .b64 param1;
	st.param.b64 	[param1], %rd13;
	.param .b64 param2;
	st.param.b64 	[param2], %rd13;
	call.uni 
	_ZN5bn25410fp_mul_oolERNS_2FpERKS0_S3_, 
	(
	param0, 
	param1, 
	param2
	);
	} // callseq 10
	cvta.to.local.u64 	%rd2, %rd13;
	ld.local.v4.u32 	{%r513, %r514, %r515, %r516}, [%rd2];
	ld.local.v4.u32 	{%r517, %r518, %r519, %r520}, [%rd2+16];
	cvta.to.local.u64 	%rd3, %rd15;
	st.local.v4.u32 	[%rd3], {%r513, %r514, %r515, %r516};
	st.local.v4.u32 	[%rd3+16], {%r517, %r518, %r519, %r520};
	{ // callseq 11, 0
	.param .b64 param0;
	st.param.b64 	[param0], %rd11;
	.param .b64 param1;
	st.param.b64 	[param1], %rd15;
	.param .b64 param2;
	st.param.b64 	[param2], %rd13;
	call.uni 
	_ZN5bn25410fp_mul_oolERNS_2FpERKS0_S3_, 
	(
	param0, 
	param1, 
