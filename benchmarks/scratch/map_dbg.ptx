.func _Z7map_dbgRN5bn2545G1AffERKNS_2FpEPS2_Pi(
	.param .b64 _Z7map_dbgRN5bn2545G1AffERKNS_2FpEPS2_Pi_param_0,
	.param .b64 _Z7map_dbgRN5bn2545G1AffERKNS_2FpEPS2_Pi_param_1,
	.param .b64 _Z7map_dbgRN5bn2545G1AffERKNS_2FpEPS2_Pi_param_2,
	.param .b64 _Z7map_dbgRN5bn2545G1AffERKNS_2FpEPS2_Pi_param_3
)
{
	.local .align 16 .b8 	__local_depot0[416];
	.reg .b64 	%SP;
	.reg .b64 	%SPL;
	.reg .pred 	%p<65>;
	.reg .b32 	%r<771>;
	.reg .b64 	%rd<108>;

	mov.u64 	%SPL, __local_depot0;
	cvta.local.u64 	%SP, %SPL;
	ld.param.u64 	%rd8, [_Z7map_dbgRN5bn2545G1AffERKNS_2FpEPS2_Pi_param_1];
	add.u64 	%rd11, %SP, 32;
	add.u64 	%rd12, %SPL, 32;
	add.u64 	%rd13, %SP, 0;
	add.u64 	%rd14, %SPL, 0;
	add.u64 	%rd15, %SP, 64;
	add.u64 	%rd16, %SPL, 64;
	add.u64 	%rd18, %SPL, 0;
	add.u64 	%rd20, %SPL, 0;
	add.u64 	%rd22, %SPL, 0;
	add.u64 	%rd24, %SPL, 32;
	add.u64 	%rd26, %SPL, 0;
	add.u64 	%rd28, %SPL, 0;
	add.u64 	%rd30, %SPL, 0;
	add.u64 	%rd32, %SPL, 0;
	add.u64 	%rd33, %SP, 96;
	add.u64 	%rd34, %SPL, 96;
	add.u64 	%rd35, %SP, 128;
	add.u64 	%rd36, %SPL, 128;
	add.u64 	%rd37, %SP, 160;
	add.u64 	%rd38, %SPL, 160;
	add.u64 	%rd39, %SP, 192;
	add.u64 	%rd40, %SPL, 192;
	add.u64 	%rd41, %SP, 224;
	add.u64 	%rd42, %SPL, 224;
	add.u64 	%rd43, %SP, 256;
	add.u64 	%rd44, %SPL, 256;
	add.u64 	%rd45, %SP, 288;
	add.u64 	%rd46, %SPL, 288;
	add.u64 	%rd47, %SP, 320;
	add.u64 	%rd48, %SPL, 320;
	add.u64 	%rd49, %SP, 352;
	add.u64 	%rd50, %SPL, 352;
	add.u64 	%rd51, %SP, 384;
	add.u64 	%rd52, %SPL, 384;
	mov.b32 	%r308, -1839054318;
	mov.b32 	%r309, 1856660074;
	mov.b32 	%r310, 290751008;
	mov.b32 	%r311, 1035942189;
	st.local.v4.u32 	[%rd40], {%r311, %r310, %r309, %r308};
	mov.b32 	%r312, 130388115;
	mov.b32 	%r313, -2031231853;
	mov.b32 	%r314, -513128441;
	mov.b32 	%r315, 1617150034;
	st.local.v4.u32 	[%rd40+16], {%r315, %r314, %r313, %r312};
	mov.b32 	%r316, 1373273600;
	mov.b32 	%r317, 1220524244;
	mov.b32 	%r318, 2089060103;
	mov.b32 	%r319, -2026525838;
	st.local.v4.u32 	[%rd42], {%r319, %r318, %r317, %r316};
	mov.b32 	%r320, 191934956;
	mov.b32 	%r321, 1219388102;
	mov.b32 	%r322, -1805845224;
	mov.b32 	%r323, 1019050996;
	st.local.v4.u32 	[%rd42+16], {%r323, %r322, %r321, %r320};
	mov.b32 	%r324, 436041239;
	mov.b32 	%r325, -459795039;
	mov.b32 	%r326, -1483068452;
	mov.b32 	%r327, -1602222031;
	st.local.v4.u32 	[%rd44], {%r327, %r326, %r325, %r324};
	mov.b32 	%r328, 96775862;
	mov.b32 	%r329, -327282872;
	mov.b32 	%r330, -1148384120;
	mov.b32 	%r331, 1431357361;
	st.local.v4.u32 	[%rd44+16], {%r331, %r330, %r329, %r328};
	{ // callseq 0, 0
	.param .b64 param0;
	st.param.b64 	[param0], %rd13;
	.param .b64 param1;
	st.param.b64 	[param1], %rd8;
	.param .b64 param2;
	st.param.b64 	[param2], %rd8;
	call.uni 
	_ZN5bn25410fp_mul_oolERNS_2FpERKS0_S3_, 
	(
	param0, 
	param1, 
	param2
	);
	} // callseq 0
	ld.local.v4.u32 	{%r332, %r333, %r334, %r335}, [%rd30];
	ld.local.v4.u32 	{%r336, %r337, %r338, %r339}, [%rd30+16];
	st.local.v4.u32 	[%rd32], {%r332, %r333, %r334, %r335};
	st.local.v4.u32 	[%rd32+16], {%r336, %r337, %r338, %r339};
	{ // callseq 1, 0
	.param .b64 param0;
	st.param.b64 	[param0], %rd13;
	.param .b64 param1;
	st.param.b64 	[param1], %rd13;
	.param .b64 param2;
	st.param.b64 	[param2], %rd39;
	call.uni 
	_ZN5bn25410fp_mul_oolERNS_2FpERKS0_S3_, 
	(
	param0, 
	param1, 
	param2
	);
	} // callseq 1
	ld.local.v4.u32 	{%r155, %r158, %r161, %r164}, [%rd28];
	ld.local.v4.u32 	{%r167, %r170, %r173, %r176}, [%rd28+16];
	mov.b32 	%r154, -980480611;
	// begin inline asm
	add.cc.u32 %r102, %r154, %r155;
	// end inline asm
	mov.b32 	%r157, -748862579;
	// begin inline asm
	addc.cc.u32 %r105, %r157, %r158;
	// end inline asm
	mov.b32 	%r160, -171504835;
	// begin inline asm
	addc.cc.u32 %r108, %r160, %r161;
	// end inline asm
	mov.b32 	%r163, 175696680;
	// begin inline asm
	addc.cc.u32 %r111, %r163, %r164;
	// end inline asm
	mov.b32 	%r166, 2021213740;
	// begin inline asm
	addc.cc.u32 %r114, %r166, %r167;
	// end inline asm
	mov.b32 	%r169, 1718526831;
	// begin inline asm
	addc.cc.u32 %r117, %r169, %r170;
	// end inline asm
	mov.b32 	%r172, -1710760145;
	// begin inline asm
	addc.cc.u32 %r120, %r172, %r173;
	// end inline asm
	mov.b32 	%r175, 235567041;
	// begin inline asm
	addc.u32 %r123, %r175, %r176;
	// end inline asm
	mov.b32 	%r283, -662897337;
	// begin inline asm
	sub.cc.u32 %r126, %r102, %r283;
	// end inline asm
	mov.b32 	%r286, 1008765974;
	// begin inline asm
	subc.cc.u32 %r129, %r105, %r286;
	// end inline asm
	mov.b32 	%r289, 1752287885;
	// begin inline asm
	subc.cc.u32 %r132, %r108, %r289;
	// end inline asm
	mov.b32 	%r292, -1753126255;
	// begin inline asm
	subc.cc.u32 %r135, %r111, %r292;
	// end inline asm
	mov.b32 	%r295, -2122229667;
	// begin inline asm
	subc.cc.u32 %r138, %r114, %r295;
	// end inline asm
	mov.b32 	%r298, -1202698826;
	// begin inline asm
	subc.cc.u32 %r141, %r117, %r298;
	// end inline asm
	mov.b32 	%r301, -516841431;
	// begin inline asm
	subc.cc.u32 %r144, %r120, %r301;
	// end inline asm
	mov.b32 	%r304, 811880050;
	// begin inline asm
	subc.cc.u32 %r147, %r123, %r304;
	// end inline asm
	mov.b32 	%r307, 0;
	// begin inline asm
	subc.u32 %r150, %r307, %r307;
	// end inline asm
	setp.eq.s32 	%p6, %r150, 0;
	selp.b32 	%r340, %r126, %r102, %p6;
	selp.b32 	%r341, %r129, %r105, %p6;
	selp.b32 	%r342, %r132, %r108, %p6;
	selp.b32 	%r343, %r135, %r111, %p6;
	selp.b32 	%r344, %r138, %r114, %p6;
	selp.b32 	%r345, %r141, %r117, %p6;
	selp.b32 	%r346, %r144, %r120, %p6;
	selp.b32 	%r347, %r147, %r123, %p6;
	st.local.v4.u32 	[%rd48], {%r340, %r341, %r342, %r343};
	st.local.v4.u32 	[%rd48+16], {%r344, %r345, %r346, %r347};
	// begin inline asm
	sub.cc.u32 %r180, %r154, %r155;
	// end inline asm
	// begin inline asm
	subc.cc.u32 %r181, %r157, %r158;
	// end inline asm
	// begin inline asm
	subc.cc.u32 %r182, %r160, %r161;
	// end inline asm
	// begin inline asm
	subc.cc.u32 %r183, %r163, %r164;
	// end inline asm
	// begin inline asm
	subc.cc.u32 %r184, %r166, %r167;
	// end inline asm
	// begin inline asm
	subc.cc.u32 %r185, %r169, %r170;
	// end inline asm
	// begin inline asm
	subc.cc.u32 %r186, %r172, %r173;
	// end inline asm
	// begin inline asm
	subc.cc.u32 %r187, %r175, %r176;
	// end inline asm
	// begin inline asm
	subc.u32 %r177, %r307, %r307;
	// end inline asm
	// begin inline asm
	{
	.reg .pred q;
	setp.ne.u32 q, %r177, 0;
	@q add.cc.u32 %r180, %r180, %r283;
	@q addc.cc.u32 %r181, %r181, %r286;
	@q addc.cc.u32 %r182, %r182, %r289;
	@q addc.cc.u32 %r183, %r183, %r292;
	@q addc.cc.u32 %r184, %r184, %r295;
	@q addc.cc.u32 %r185, %r185, %r298;
	@q addc.cc.u32 %r186, %r186, %r301;
	@q addc.u32 %r187, %r187, %r304;
	}
	// end inline asm
	st.local.v4.u32 	[%rd46], {%r180, %r181, %r182, %r183};
	st.local.v4.u32 	[%rd46+16], {%r184, %r185, %r186, %r187};
	{ // callseq 2, 0
	.param .b64 param0;
	st.param.b64 	[param0], %rd13;
	.param .b64 param1;
	st.param.b64 	[param1], %rd45;
	.param .b64 param2;
	st.param.b64 	[param2], %rd47;
	call.uni 
	_ZN5bn25410fp_mul_oolERNS_2FpERKS0_S3_, 
	(
	param0, 
	param1, 
	param2
	);
	} // callseq 2
	ld.local.v4.u32 	{%r348, %r349, %r350, %r351}, [%rd26];
	ld.local.v4.u32 	{%r352, %r353, %r354, %r355}, [%rd26+16];
	st.local.v4.u32 	[%rd34], {%r348, %r349, %r350, %r351};
	st.local.v4.u32 	[%rd34+16], {%r352, %r353, %r354, %r355};
	{ // callseq 3, 0
	.param .b64 param0;
	st.param.b64 	[param0], %rd11;
	.param .b64 param1;
	st.param.b64 	[param1], %rd33;
	call.uni 
	_ZN5bn25410fp_inv_oolERNS_2FpERKS0_, 
	(
	param0, 
	param1
	);
	} // callseq 3
	ld.local.v4.u32 	{%r356, %r357, %r358, %r359}, [%rd24];
	ld.local.v4.u32 	{%r360, %r361, %r362, %r363}, [%rd24+16];
	st.local.v4.u32 	[%rd50], {%r356, %r357, %r358, %r359};
	st.local.v4.u32 	[%rd50+16], {%r360, %r361, %r362, %r363};
	{ // callseq 4, 0
	.param .b64 param0;
	st.param.b64 	[param0], %rd13;
	.param .b64 param1;
	st.param.b64 	[param1], %rd8;
	.param .b64 param2;
	st.param.b64 	[param2], %rd45;
	call.uni 
	_ZN5bn25410fp_mul_oolERNS_2FpERKS0_S3_, 
	(
	param0, 
	param1, 
	param2
	);
	} // callseq 4
	ld.local.v4.u32 	{%r364, %r365, %r366, %r367}, [%rd22];
	ld.local.v4.u32 	{%r368, %r369, %r370, %r371}, [%rd22+16];
	st.local.v4.u32 	[%rd38], {%r364, %r365, %r366, %r367};
	st.local.v4.u32 	[%rd38+16], {%r368, %r369, %r370, %r371};
	{ // callseq 5, 0
	.param .b64 param0;
	st.param.b64 	[param0], %rd13;
	.param .b64 param1;
	st.param.b64 	[param1], %rd37;
	.param .b64 param2;
	st.param.b64 	[param2], %rd49;
	call.uni 
	_ZN5bn25410fp_mul_oolERNS_2FpERKS0_S3_, 
	(
	param0, 
	param1, 
	param2
	);
	} // callseq 5
	ld.local.v4.u32 	{%r372, %r373, %r374, %r375}, [%rd20];
	ld.local.v4.u32 	{%r376, %r377, %r378, %r379}, [%rd20+16];
	st.local.v4.u32 	[%rd36], {%r372, %r373, %r374, %r375};
	st.local.v4.u32 	[%rd36+16], {%r376, %r377, %r378, %r379};
	{ // callseq 6, 0
	.param .b64 param0;
	st.param.b64 	[param0], %rd13;
	.param .b64 param1;
	st.param.b64 	[param1], %rd35;
	.param .b64 param2;
	st.param.b64 	[param2], %rd41;
	call.uni 
	_ZN5bn25410fp_mul_oolERNS_2FpERKS0_S3_, 
	(
	param0, 
	param1, 
	param2
	);
	} // callseq 6
	ld.local.v4.u32 	{%r207, %r210, %r213, %r216}, [%rd18];
	ld.local.v4.u32 	{%r219, %r222, %r225, %r228}, [%rd18+16];
	mov.b32 	%r206, -1988692011;
	// begin inline asm
	sub.cc.u32 %r232, %r206, %r207;
	// end inline asm
	mov.b32 	%r209, -1268669372;
	// begin inline asm
	subc.cc.u32 %r233, %r209, %r210;
	// end inline asm
	mov.b32 	%r212, 961896359;
	// begin inline asm
	subc.cc.u32 %r234, %r212, %r213;
	// end inline asm
	mov.b32 	%r215, -964411468;
	// begin inline asm
	subc.cc.u32 %r235, %r215, %r216;
	// end inline asm
	mov.b32 	%r218, -2071721704;
	// begin inline asm
	subc.cc.u32 %r236, %r218, %r219;
	// end inline asm
	mov.b32 	%r221, 686870819;
	// begin inline asm
	subc.cc.u32 %r237, %r221, %r222;
	// end inline asm
	mov.b32 	%r224, -1550524291;
	// begin inline asm
	subc.cc.u32 %r238, %r224, %r225;
	// end inline asm
	mov.b32 	%r227, 288156504;
	// begin inline asm
	subc.cc.u32 %r239, %r227, %r228;
	// end inline asm
	// begin inline asm
	subc.u32 %r229, %r307, %r307;
	// end inline asm
	// begin inline asm
	{
	.reg .pred q;
	setp.ne.u32 q, %r229, 0;
	@q add.cc.u32 %r232, %r232, %r283;
	@q addc.cc.u32 %r233, %r233, %r286;
	@q addc.cc.u32 %r234, %r234, %r289;
	@q addc.cc.u32 %r235, %r235, %r292;
	@q addc.cc.u32 %r236, %r236, %r295;
	@q addc.cc.u32 %r237, %r237, %r298;
	@q addc.cc.u32 %r238, %r238, %r301;
	@q addc.u32 %r239, %r239, %r304;
	}
	// end inline asm
	st.local.v4.u32 	[%rd52], {%r232, %r233, %r234, %r235};
	st.local.v4.u32 	[%rd52+16], {%r236, %r237, %r238, %r239};
	{ // callseq 7, 0
	.param .b64 param0;
	st.param.b64 	[param0], %rd13;
	.param .b64 param1;
	st.param.b64 	[param1], %rd51;
	.param .b64 param2;
	st.param.b64 	[param2], %rd51;
	call.uni 
	_ZN5bn25410fp_mul_oolERNS_2FpERKS0_S3_, 
	(
	param0, 
	param1, 
	param2
	);
	} // callseq 7
	ld.local.v4.u32 	{%r380, %r381, %r382, %r383}, [%rd14];
	ld.local.v4.u32 	{%r384, %r385, %r386, %r387}, [%rd14+16];
	st.local.v4.u32 	[%rd16], {%r380, %r381, %r382, %r383};
	st.local.v4.u32 	[%rd16+16], {%r384, %r385, %r386, %r387};
	{ // callseq 8, 0
	.param .b64 param0;
	st.param.b64 	[param0], %rd11;
	.param .b64 param1;
	st.param.b64 	[param1], %rd15;
	.param .b64 param2;
	st.param.b64 	[param2], %rd51;
	call.uni 
	_ZN5bn25410fp_mul_oolERNS_2FpERKS0_S3_, 
	(
	param0, 
	param1, 
	param2
	);
	} // callseq 8
	ld.local.v4.u32 	{%r258, %r261, %r264, %r267}, [%rd12];
	ld.local.v4.u32 	{%r270, %r273, %r276, %r279}, [%rd12+16];
	mov.b32 	%r259, 1353525463;
	// begin inline asm
	add.cc.u32 %r257, %r258, %r259;
	// end inline asm
	mov.b32 	%r262, 2048379561;
	// begin inline asm
	addc.cc.u32 %r260, %r261, %r262;
	// end inline asm
	mov.b32 	%r265, -514514503;
	// begin inline asm
	addc.cc.u32 %r263, %r264, %r265;
	// end inline asm
	mov.b32 	%r268, 527090042;
	// begin inline asm
	addc.cc.u32 %r266, %r267, %r268;
	// end inline asm
	mov.b32 	%r271, 1768673924;
	// begin inline asm
	addc.cc.u32 %r269, %r270, %r271;
	// end inline asm
	mov.b32 	%r274, 860613198;
	// begin inline asm
	addc.cc.u32 %r272, %r273, %r274;
	// end inline asm
	mov.b32 	%r277, -837313138;
	// begin inline asm
	addc.cc.u32 %r275, %r276, %r277;
	// end inline asm
	mov.b32 	%r280, 706701124;
	// begin inline asm
	addc.u32 %r278, %r279, %r280;
	// end inline asm
	// begin inline asm
	sub.cc.u32 %r281, %r257, %r283;
	// end inline asm
	// begin inline asm
	subc.cc.u32 %r284, %r260, %r286;
	// end inline asm
	// begin inline asm
	subc.cc.u32 %r287, %r263, %r289;
	// end inline asm
	// begin inline asm
	subc.cc.u32 %r290, %r266, %r292;
	// end inline asm
	// begin inline asm
	subc.cc.u32 %r293, %r269, %r295;
	// end inline asm
	// begin inline asm
	subc.cc.u32 %r296, %r272, %r298;
	// end inline asm
	// begin inline asm
	subc.cc.u32 %r299, %r275, %r301;
	// end inline asm
	// begin inline asm
	subc.cc.u32 %r302, %r278, %r304;
	// end inline asm
	// begin inline asm
	subc.u32 %r305, %r307, %r307;
	// end inline asm
	setp.eq.s32 	%p7, %r305, 0;
	selp.b32 	%r388, %r281, %r257, %p7;
	selp.b32 	%r389, %r284, %r260, %p7;
	selp.b32 	%r390, %r287, %r263, %p7;
	selp.b32 	%r391, %r290, %r266, %p7;
	selp.b32 	%r392, %r293, %r269, %p7;
	selp.b32 	%r393, %r296, %r272, %p7;
	selp.b32 	%r394, %r299, %r275, %p7;
	selp.b32 	%r395, %r302, %r278, %p7;
	mov.u64 	%rd53, _ZN41_INTERNAL_be76c628_10_dbg_map_cu_c47dc1785bn2547FP_PM1HE;
	cvta.const.u64 	%rd54, %rd53;
	{ // callseq 9, 0
	.param .b64 param0;
	st.param.b64 	[param0], %rd13;
	.param .align 16 .b8 param1[32];
	st.param.v4.b32 	[param1], {%r388, %r389, %r390, %r391};
	st.param.v4.b32 	[param1+16], {%r392, %r393, %r394, %r395};
	.param .b64 param2;
	st.param.b64 	[param2], %rd54;
	call.uni 
	_ZN5bn25412fp_pow_fixedERNS_2FpES0_PKj, 
	(
	param0, 
	param1, 
	param2
	);
	} // callseq 9
	or.b32  	%r396, %r389, %r388;
	or.b32  	%r397, %r390, %r396;
	or.b32  	%r398, %r391, %r397;
	or.b32  	%r399, %r392, %r398;
	or.b32  	%r400, %r393, %r399;
	or.b32  	%r401, %r394, %r400;
	or.b32  	%r402, %r395, %r401;
	setp.eq.s32 	%p8, %r402, 0;
	mov.pred 	%p63, -1;
	@%p8 bra 	$L__BB0_2;
	cvta.to.local.u64 	%rd57, %rd13;
	ld.local.v4.u32 	{%r403, %r404, %r405, %r406}, [%rd57];
	ld.local.v4.u32 	{%r407, %r408, %r409, %r410}, [%rd57+16];
	setp.eq.s32 	%p9, %r404, -748862579;
	setp.eq.s32 	%p10, %r403, -980480611;
	and.pred  	%p11, %p9, %p10;
	setp.eq.s32 	%p12, %r405, -171504835;
	and.pred  	%p13, %p11, %p12;
	setp.eq.s32 	%p14, %r406, 175696680;
	and.pred  	%p15, %p13, %p14;
	setp.eq.s32 	%p16, %r407, 2021213740;
	and.pred  	%p17, %p15, %p16;
	setp.eq.s32 	%p18, %r408, 1718526831;
	and.pred  	%p19, %p17, %p18;
	setp.eq.s32 	%p20, %r409, -1710760145;
	and.pred  	%p21, %p19, %p20;
	setp.eq.s32 	%p22, %r410, 235567041;
	and.pred  	%p63, %p21, %p22;
$L__BB0_2:
	ld.param.u64 	%rd106, [_Z7map_dbgRN5bn2545G1AffERKNS_2FpEPS2_Pi_param_2];
	mov.b32 	%r412, -1988692011;
	// begin inline asm
	add.cc.u32 %r411, %r412, %r207;
	// end inline asm
	mov.b32 	%r415, -1268669372;
	// begin inline asm
	addc.cc.u32 %r414, %r415, %r210;
	// end inline asm
	mov.b32 	%r418, 961896359;
	// begin inline asm
	addc.cc.u32 %r417, %r418, %r213;
	// end inline asm
	mov.b32 	%r421, -964411468;
	// begin inline asm
	addc.cc.u32 %r420, %r421, %r216;
	// end inline asm
	mov.b32 	%r424, -2071721704;
	// begin inline asm
	addc.cc.u32 %r423, %r424, %r219;
	// end inline asm
	mov.b32 	%r427, 686870819;
	// begin inline asm
	addc.cc.u32 %r426, %r427, %r222;
	// end inline asm
	mov.b32 	%r430, -1550524291;
	// begin inline asm
	addc.cc.u32 %r429, %r430, %r225;
	// end inline asm
	mov.b32 	%r433, 288156504;
	// begin inline asm
	addc.u32 %r432, %r433, %r228;
	// end inline asm
	mov.b32 	%r488, -662897337;
	// begin inline asm
	sub.cc.u32 %r435, %r411, %r488;
	// end inline asm
	mov.b32 	%r491, 1008765974;
	// begin inline asm
	subc.cc.u32 %r438, %r414, %r491;
	// end inline asm
	mov.b32 	%r494, 1752287885;
	// begin inline asm
	subc.cc.u32 %r441, %r417, %r494;
	// end inline asm
	mov.b32 	%r497, -1753126255;
	// begin inline asm
	subc.cc.u32 %r444, %r420, %r497;
	// end inline asm
	mov.b32 	%r500, -2122229667;
	// begin inline asm
	subc.cc.u32 %r447, %r423, %r500;
	// end inline asm
	mov.b32 	%r503, -1202698826;
	// begin inline asm
	subc.cc.u32 %r450, %r426, %r503;
	// end inline asm
	mov.b32 	%r506, -516841431;
	// begin inline asm
	subc.cc.u32 %r453, %r429, %r506;
	// end inline asm
	mov.b32 	%r509, 811880050;
	// begin inline asm
	subc.cc.u32 %r456, %r432, %r509;
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
	param2
	);
	} // callseq 11
	cvta.to.local.u64 	%rd4, %rd11;
	ld.local.v4.u32 	{%r463, %r466, %r469, %r472}, [%rd4];
	ld.local.v4.u32 	{%r475, %r478, %r481, %r484}, [%rd4+16];
	mov.b32 	%r464, 1353525463;
	// begin inline asm
	add.cc.u32 %r462, %r463, %r464;
	// end inline asm
	mov.b32 	%r467, 2048379561;
	// begin inline asm
	addc.cc.u32 %r465, %r466, %r467;
	// end inline asm
	mov.b32 	%r470, -514514503;
	// begin inline asm
	addc.cc.u32 %r468, %r469, %r470;
	// end inline asm
	mov.b32 	%r473, 527090042;
	// begin inline asm
	addc.cc.u32 %r471, %r472, %r473;
	// end inline asm
	mov.b32 	%r476, 1768673924;
	// begin inline asm
	addc.cc.u32 %r474, %r475, %r476;
	// end inline asm
	mov.b32 	%r479, 860613198;
	// begin inline asm
	addc.cc.u32 %r477, %r478, %r479;
	// end inline asm
	mov.b32 	%r482, -837313138;
	// begin inline asm
	addc.cc.u32 %r480, %r481, %r482;
	// end inline asm
	mov.b32 	%r485, 706701124;
	// begin inline asm
	addc.u32 %r483, %r484, %r485;
	// end inline asm
	// begin inline asm
	sub.cc.u32 %r486, %r462, %r488;
	// end inline asm
	// begin inline asm
	subc.cc.u32 %r489, %r465, %r491;
	// end inline asm
	// begin inline asm
	subc.cc.u32 %r492, %r468, %r494;
	// end inline asm
	// begin inline asm
	subc.cc.u32 %r495, %r471, %r497;
	// end inline asm
	// begin inline asm
	subc.cc.u32 %r498, %r474, %r500;
	// end inline asm
	// begin inline asm
	subc.cc.u32 %r501, %r477, %r503;
	// end inline asm
	// begin inline asm
	subc.cc.u32 %r504, %r480, %r506;
	// end inline asm
	// begin inline asm
	subc.cc.u32 %r507, %r483, %r509;
	// end inline asm
	// begin inline asm
	subc.u32 %r510, %r512, %r512;
	// end inline asm
	setp.eq.s32 	%p25, %r510, 0;
	selp.b32 	%r521, %r486, %r462, %p25;
	selp.b32 	%r522, %r489, %r465, %p25;
	selp.b32 	%r523, %r492, %r468, %p25;
	selp.b32 	%r524, %r495, %r471, %p25;
	selp.b32 	%r525, %r498, %r474, %p25;
	selp.b32 	%r526, %r501, %r477, %p25;
	selp.b32 	%r527, %r504, %r480, %p25;
	selp.b32 	%r528, %r507, %r483, %p25;
	cvta.const.u64 	%rd63, %rd53;
	{ // callseq 12, 0
	.param .b64 param0;
	st.param.b64 	[param0], %rd13;
	.param .align 16 .b8 param1[32];
	st.param.v4.b32 	[param1], {%r521, %r522, %r523, %r524};
	st.param.v4.b32 	[param1+16], {%r525, %r526, %r527, %r528};
	.param .b64 param2;
	st.param.b64 	[param2], %rd63;
	call.uni 
	_ZN5bn25412fp_pow_fixedERNS_2FpES0_PKj, 
	(
	param0, 
	param1, 
	param2
	);
	} // callseq 12
	cvta.to.global.u64 	%rd5, %rd106;
	st.global.v4.u32 	[%rd5+256], {%r521, %r522, %r523, %r524};
	st.global.v4.u32 	[%rd5+272], {%r525, %r526, %r527, %r528};
	cvta.to.local.u64 	%rd6, %rd13;
	ld.local.v4.u32 	{%r31, %r32, %r30, %r29}, [%rd6];
	ld.local.v4.u32 	{%r36, %r35, %r34, %r33}, [%rd6+16];
	st.global.v4.u32 	[%rd5+288], {%r31, %r32, %r30, %r29};
	st.global.v4.u32 	[%rd5+304], {%r36, %r35, %r34, %r33};
	{ // callseq 13, 0
	.param .b64 param0;
	st.param.b64 	[param0], %rd13;
	.param .align 16 .b8 param1[32];
	st.param.v4.b32 	[param1], {%r521, %r522, %r523, %r524};
	st.param.v4.b32 	[param1+16], {%r525, %r526, %r527, %r528};
	.param .b64 param2;
	st.param.b64 	[param2], %rd63;
	call.uni 
	_ZN5bn25412fp_pow_fixedERNS_2FpES0_PKj, 
	(
	param0, 
	param1, 
	param2
	);
	} // callseq 13
	or.b32  	%r529, %r522, %r521;
	or.b32  	%r530, %r523, %r529;
	or.b32  	%r531, %r524, %r530;
	or.b32  	%r532, %r525, %r531;
	or.b32  	%r533, %r526, %r532;
	or.b32  	%r534, %r527, %r533;
	or.b32  	%r37, %r528, %r534;
	setp.eq.s32 	%p26, %r37, 0;
	mov.pred 	%p64, -1;
	@%p26 bra 	$L__BB0_4;
	cvta.to.local.u64 	%rd67, %rd13;
	ld.local.v4.u32 	{%r535, %r536, %r537, %r538}, [%rd67];
	ld.local.v4.u32 	{%r539, %r540, %r541, %r542}, [%rd67+16];
	setp.eq.s32 	%p27, %r536, -748862579;
	setp.eq.s32 	%p28, %r535, -980480611;
	and.pred  	%p29, %p27, %p28;
	setp.eq.s32 	%p30, %r537, -171504835;
	and.pred  	%p31, %p29, %p30;
	setp.eq.s32 	%p32, %r538, 175696680;
	and.pred  	%p33, %p31, %p32;
	setp.eq.s32 	%p34, %r539, 2021213740;
	and.pred  	%p35, %p33, %p34;
	setp.eq.s32 	%p36, %r540, 1718526831;
	and.pred  	%p37, %p35, %p36;
	setp.eq.s32 	%p38, %r541, -1710760145;
	and.pred  	%p39, %p37, %p38;
	setp.eq.s32 	%p40, %r542, 235567041;
	and.pred  	%p64, %p39, %p40;
$L__BB0_4:
	ld.param.u64 	%rd107, [_Z7map_dbgRN5bn2545G1AffERKNS_2FpEPS2_Pi_param_3];
	ld.param.u64 	%rd105, [_Z7map_dbgRN5bn2545G1AffERKNS_2FpEPS2_Pi_param_1];
	cvta.to.global.u64 	%rd68, %rd107;
	setp.eq.s32 	%p41, %r37, 0;
	selp.u32 	%r645, 1, 0, %p64;
	st.global.u32 	[%rd68+12], %r645;
	setp.eq.s32 	%p42, %r32, -748862579;
	setp.eq.s32 	%p43, %r31, -980480611;
	and.pred  	%p44, %p42, %p43;
	setp.eq.s32 	%p45, %r30, -171504835;
	and.pred  	%p46, %p44, %p45;
	setp.eq.s32 	%p47, %r29, 175696680;
	and.pred  	%p48, %p46, %p47;
	setp.eq.s32 	%p49, %r36, 2021213740;
	and.pred  	%p50, %p48, %p49;
	setp.eq.s32 	%p51, %r35, 1718526831;
	and.pred  	%p52, %p50, %p51;
	setp.eq.s32 	%p53, %r34, -1710760145;
	and.pred  	%p54, %p52, %p53;
	setp.eq.s32 	%p55, %r33, 235567041;
	and.pred  	%p56, %p54, %p55;
	selp.u32 	%r646, 1, 0, %p56;
	st.global.u32 	[%rd68+16], %r646;
	selp.u32 	%r647, 1, 0, %p41;
	st.global.u32 	[%rd68+20], %r647;
	not.pred 	%p57, %p63;
	and.pred  	%p58, %p64, %p57;
	{ // callseq 14, 0
	.param .b64 param0;
	st.param.b64 	[param0], %rd13;
	.param .b64 param1;
	st.param.b64 	[param1], %rd47;
	.param .b64 param2;
	st.param.b64 	[param2], %rd47;
	call.uni 
	_ZN5bn25410fp_mul_oolERNS_2FpERKS0_S3_, 
	(
	param0, 
	param1, 
	param2
	);
	} // callseq 14
	cvta.to.local.u64 	%rd71, %rd13;
	ld.local.v4.u32 	{%r648, %r649, %r650, %r651}, [%rd71];
	ld.local.v4.u32 	{%r652, %r653, %r654, %r655}, [%rd71+16];
	cvta.to.local.u64 	%rd73, %rd33;
	st.local.v4.u32 	[%rd73], {%r648, %r649, %r650, %r651};
	st.local.v4.u32 	[%rd73+16], {%r652, %r653, %r654, %r655};
	{ // callseq 15, 0
	.param .b64 param0;
	st.param.b64 	[param0], %rd13;
	.param .b64 param1;
	st.param.b64 	[param1], %rd33;
	.param .b64 param2;
	st.param.b64 	[param2], %rd49;
	call.uni 
	_ZN5bn25410fp_mul_oolERNS_2FpERKS0_S3_, 
	(
	param0, 
	param1, 
	param2
	);
	} // callseq 15
	cvta.to.local.u64 	%rd76, %rd13;
	ld.local.v4.u32 	{%r656, %r657, %r658, %r659}, [%rd76];
	ld.local.v4.u32 	{%r660, %r661, %r662, %r663}, [%rd76+16];
	cvta.to.local.u64 	%rd78, %rd37;
	st.local.v4.u32 	[%rd78], {%r656, %r657, %r658, %r659};
	st.local.v4.u32 	[%rd78+16], {%r660, %r661, %r662, %r663};
	{ // callseq 16, 0
	.param .b64 param0;
	st.param.b64 	[param0], %rd13;
	.param .b64 param1;
	st.param.b64 	[param1], %rd37;
	.param .b64 param2;
	st.param.b64 	[param2], %rd37;
	call.uni 
	_ZN5bn25410fp_mul_oolERNS_2FpERKS0_S3_, 
	(
	param0, 
	param1, 
	param2
	);
	} // callseq 16
	cvta.to.local.u64 	%rd80, %rd13;
	ld.local.v4.u32 	{%r664, %r665, %r666, %r667}, [%rd80];
	ld.local.v4.u32 	{%r668, %r669, %r670, %r671}, [%rd80+16];
	cvta.to.local.u64 	%rd82, %rd35;
	st.local.v4.u32 	[%rd82], {%r664, %r665, %r666, %r667};
	st.local.v4.u32 	[%rd82+16], {%r668, %r669, %r670, %r671};
	{ // callseq 17, 0
	.param .b64 param0;
	st.param.b64 	[param0], %rd13;
	.param .b64 param1;
	st.param.b64 	[param1], %rd35;
	.param .b64 param2;
	st.param.b64 	[param2], %rd43;
	call.uni 
	_ZN5bn25410fp_mul_oolERNS_2FpERKS0_S3_, 
	(
	param0, 
	param1, 
	param2
	);
	} // callseq 17
	cvta.to.local.u64 	%rd85, %rd13;
	ld.local.v4.u32 	{%r544, %r547, %r550, %r553}, [%rd85];
	ld.local.v4.u32 	{%r556, %r559, %r562, %r565}, [%rd85+16];
	mov.b32 	%r545, -980480611;
	// begin inline asm
	add.cc.u32 %r543, %r544, %r545;
	// end inline asm
	mov.b32 	%r548, -748862579;
	// begin inline asm
	addc.cc.u32 %r546, %r547, %r548;
	// end inline asm
	mov.b32 	%r551, -171504835;
	// begin inline asm
	addc.cc.u32 %r549, %r550, %r551;
	// end inline asm
	mov.b32 	%r554, 175696680;
	// begin inline asm
	addc.cc.u32 %r552, %r553, %r554;
	// end inline asm
	mov.b32 	%r557, 2021213740;
	// begin inline asm
	addc.cc.u32 %r555, %r556, %r557;
	// end inline asm
	mov.b32 	%r560, 1718526831;
	// begin inline asm
	addc.cc.u32 %r558, %r559, %r560;
	// end inline asm
	mov.b32 	%r563, -1710760145;
	// begin inline asm
	addc.cc.u32 %r561, %r562, %r563;
	// end inline asm
	mov.b32 	%r566, 235567041;
	// begin inline asm
	addc.u32 %r564, %r565, %r566;
	// end inline asm
	mov.b32 	%r620, -662897337;
	// begin inline asm
	sub.cc.u32 %r567, %r543, %r620;
	// end inline asm
	mov.b32 	%r623, 1008765974;
	// begin inline asm
	subc.cc.u32 %r570, %r546, %r623;
	// end inline asm
	mov.b32 	%r626, 1752287885;
	// begin inline asm
	subc.cc.u32 %r573, %r549, %r626;
	// end inline asm
	mov.b32 	%r629, -1753126255;
	// begin inline asm
	subc.cc.u32 %r576, %r552, %r629;
	// end inline asm
	mov.b32 	%r632, -2122229667;
	// begin inline asm
	subc.cc.u32 %r579, %r555, %r632;
	// end inline asm
	mov.b32 	%r635, -1202698826;
	// begin inline asm
	subc.cc.u32 %r582, %r558, %r635;
	// end inline asm
	mov.b32 	%r638, -516841431;
	// begin inline asm
	subc.cc.u32 %r585, %r561, %r638;
	// end inline asm
	mov.b32 	%r641, 811880050;
	// begin inline asm
	subc.cc.u32 %r588, %r564, %r641;
	// end inline asm
	mov.b32 	%r644, 0;
	// begin inline asm
	subc.u32 %r591, %r644, %r644;
	// end inline asm
	setp.eq.s32 	%p59, %r591, 0;
	selp.b32 	%r672, %r567, %r543, %p59;
	selp.b32 	%r673, %r570, %r546, %p59;
	selp.b32 	%r674, %r573, %r549, %p59;
	selp.b32 	%r675, %r576, %r552, %p59;
	selp.b32 	%r676, %r579, %r555, %p59;
	selp.b32 	%r677, %r582, %r558, %p59;
	selp.b32 	%r678, %r585, %r561, %p59;
	selp.b32 	%r679, %r588, %r564, %p59;
	st.local.v4.u32 	[%rd78], {%r672, %r673, %r674, %r675};
	st.local.v4.u32 	[%rd78+16], {%r676, %r677, %r678, %r679};
	selp.u32 	%r680, 1, 0, %p63;
	st.global.u32 	[%rd68+24], %r680;
	selp.u32 	%r681, 1, 0, %p58;
	st.global.u32 	[%rd68+28], %r681;
	st.global.v4.u32 	[%rd5], {%r232, %r233, %r234, %r235};
	st.global.v4.u32 	[%rd5+16], {%r236, %r237, %r238, %r239};
	st.global.v4.u32 	[%rd5+32], {%r17, %r18, %r19, %r20};
	st.global.v4.u32 	[%rd5+48], {%r25, %r26, %r23, %r24};
	st.global.v4.u32 	[%rd5+64], {%r672, %r673, %r674, %r675};
	st.global.v4.u32 	[%rd5+80], {%r676, %r677, %r678, %r679};
	selp.b32 	%r682, %r232, %r672, %p63;
	selp.b32 	%r683, %r233, %r673, %p63;
	selp.b32 	%r684, %r234, %r674, %p63;
	selp.b32 	%r685, %r235, %r675, %p63;
	selp.b32 	%r686, %r236, %r676, %p63;
	selp.b32 	%r687, %r237, %r677, %p63;
	selp.b32 	%r688, %r238, %r678, %p63;
	selp.b32 	%r689, %r239, %r679, %p63;
	st.global.v4.u32 	[%rd5+96], {%r682, %r683, %r684, %r685};
	st.global.v4.u32 	[%rd5+112], {%r686, %r687, %r688, %r689};
	selp.b32 	%r58, %r17, %r682, %p58;
	selp.b32 	%r59, %r18, %r683, %p58;
	selp.b32 	%r60, %r19, %r684, %p58;
	selp.b32 	%r61, %r20, %r685, %p58;
	selp.b32 	%r74, %r25, %r686, %p58;
	selp.b32 	%r75, %r26, %r687, %p58;
	selp.b32 	%r76, %r23, %r688, %p58;
	selp.b32 	%r77, %r24, %r689, %p58;
	cvta.to.local.u64 	%rd87, %rd39;
	st.local.v4.u32 	[%rd87], {%r58, %r59, %r60, %r61};
	st.local.v4.u32 	[%rd87+16], {%r74, %r75, %r76, %r77};
	st.global.v4.u32 	[%rd5+128], {%r58, %r59, %r60, %r61};
	st.global.v4.u32 	[%rd5+144], {%r74, %r75, %r76, %r77};
	{ // callseq 18, 0
	.param .b64 param0;
	st.param.b64 	[param0], %rd13;
	.param .b64 param1;
	st.param.b64 	[param1], %rd39;
	.param .b64 param2;
	st.param.b64 	[param2], %rd39;
	call.uni 
	_ZN5bn25410fp_mul_oolERNS_2FpERKS0_S3_, 
	(
	param0, 
	param1, 
	param2
	);
	} // callseq 18
	cvta.to.local.u64 	%rd89, %rd13;
	ld.local.v4.u32 	{%r690, %r691, %r692, %r693}, [%rd89];
	ld.local.v4.u32 	{%r694, %r695, %r696, %r697}, [%rd89+16];
	cvta.to.local.u64 	%rd91, %rd11;
	st.local.v4.u32 	[%rd91], {%r690, %r691, %r692, %r693};
	st.local.v4.u32 	[%rd91+16], {%r694, %r695, %r696, %r697};
	{ // callseq 19, 0
	.param .b64 param0;
	st.param.b64 	[param0], %rd13;
	.param .b64 param1;
	st.param.b64 	[param1], %rd11;
	.param .b64 param2;
	st.param.b64 	[param2], %rd39;
	call.uni 
	_ZN5bn25410fp_mul_oolERNS_2FpERKS0_S3_, 
	(
	param0, 
	param1, 
	param2
	);
	} // callseq 19
	ld.local.v4.u32 	{%r595, %r598, %r601, %r604}, [%rd6];
	ld.local.v4.u32 	{%r607, %r610, %r613, %r616}, [%rd6+16];
	mov.b32 	%r596, 1353525463;
	// begin inline asm
	add.cc.u32 %r594, %r595, %r596;
	// end inline asm
	mov.b32 	%r599, 2048379561;
	// begin inline asm
	addc.cc.u32 %r597, %r598, %r599;
	// end inline asm
	mov.b32 	%r602, -514514503;
	// begin inline asm
	addc.cc.u32 %r600, %r601, %r602;
	// end inline asm
	mov.b32 	%r605, 527090042;
	// begin inline asm
	addc.cc.u32 %r603, %r604, %r605;
	// end inline asm
	mov.b32 	%r608, 1768673924;
	// begin inline asm
	addc.cc.u32 %r606, %r607, %r608;
	// end inline asm
	mov.b32 	%r611, 860613198;
	// begin inline asm
	addc.cc.u32 %r609, %r610, %r611;
	// end inline asm
	mov.b32 	%r614, -837313138;
	// begin inline asm
	addc.cc.u32 %r612, %r613, %r614;
	// end inline asm
	mov.b32 	%r617, 706701124;
	// begin inline asm
	addc.u32 %r615, %r616, %r617;
	// end inline asm
	// begin inline asm
	sub.cc.u32 %r618, %r594, %r620;
	// end inline asm
	// begin inline asm
	subc.cc.u32 %r621, %r597, %r623;
	// end inline asm
	// begin inline asm
	subc.cc.u32 %r624, %r600, %r626;
	// end inline asm
	// begin inline asm
	subc.cc.u32 %r627, %r603, %r629;
	// end inline asm
	// begin inline asm
	subc.cc.u32 %r630, %r606, %r632;
	// end inline asm
	// begin inline asm
	subc.cc.u32 %r633, %r609, %r635;
	// end inline asm
	// begin inline asm
	subc.cc.u32 %r636, %r612, %r638;
	// end inline asm
	// begin inline asm
	subc.cc.u32 %r639, %r615, %r641;
	// end inline asm
	// begin inline asm
	subc.u32 %r642, %r644, %r644;
	// end inline asm
	setp.eq.s32 	%p60, %r642, 0;
	selp.b32 	%r698, %r618, %r594, %p60;
	selp.b32 	%r699, %r621, %r597, %p60;
	selp.b32 	%r700, %r624, %r600, %p60;
	selp.b32 	%r701, %r627, %r603, %p60;
	selp.b32 	%r702, %r630, %r606, %p60;
	selp.b32 	%r703, %r633, %r609, %p60;
	selp.b32 	%r704, %r636, %r612, %p60;
	selp.b32 	%r705, %r639, %r615, %p60;
	mov.u64 	%rd93, _ZN41_INTERNAL_be76c628_10_dbg_map_cu_c47dc1785bn2547FP_PP1QE;
	cvta.const.u64 	%rd94, %rd93;
	{ // callseq 20, 0
	.param .b64 param0;
	st.param.b64 	[param0], %rd13;
	.param .align 16 .b8 param1[32];
	st.param.v4.b32 	[param1], {%r698, %r699, %r700, %r701};
	st.param.v4.b32 	[param1+16], {%r702, %r703, %r704, %r705};
	.param .b64 param2;
	st.param.b64 	[param2], %rd94;
	call.uni 
	_ZN5bn25412fp_pow_fixedERNS_2FpES0_PKj, 
	(
	param0, 
	param1, 
	param2
	);
	} // callseq 20
	ld.local.v4.u32 	{%r770, %r769, %r768, %r767}, [%rd1];
	ld.local.v4.u32 	{%r766, %r765, %r764, %r763}, [%rd1+16];
	cvta.to.local.u64 	%rd97, %rd41;
	st.local.v4.u32 	[%rd97], {%r770, %r769, %r768, %r767};
	st.local.v4.u32 	[%rd97+16], {%r766, %r765, %r764, %r763};
	mov.b32 	%r706, 1;
	st.local.v4.u32 	[%rd3], {%r706, %r644, %r644, %r644};
	st.local.v4.u32 	[%rd3+16], {%r644, %r644, %r644, %r644};
	{ // callseq 21, 0
	.param .b64 param0;
	st.param.b64 	[param0], %rd13;
	.param .b64 param1;
	st.param.b64 	[param1], %rd105;
	.param .b64 param2;
	st.param.b64 	[param2], %rd15;
	call.uni 
	_ZN5bn25410fp_mul_oolERNS_2FpERKS0_S3_, 
	(
	param0, 
	param1, 
	param2
	);
	} // callseq 21
	ld.local.u32 	%r707, [%rd2];
	st.local.v4.u32 	[%rd4], {%r706, %r644, %r644, %r644};
	st.local.v4.u32 	[%rd4+16], {%r644, %r644, %r644, %r644};
	{ // callseq 22, 0
	.param .b64 param0;
	st.param.b64 	[param0], %rd13;
	.param .b64 param1;
	st.param.b64 	[param1], %rd41;
	.param .b64 param2;
	st.param.b64 	[param2], %rd11;
	call.uni 
	_ZN5bn25410fp_mul_oolERNS_2FpERKS0_S3_, 
	(
	param0, 
	param1, 
	param2
	);
	} // callseq 22
	cvta.to.local.u64 	%rd102, %rd13;
	ld.local.u32 	%r708, [%rd102];
	xor.b32  	%r709, %r708, %r707;
	and.b32  	%r710, %r709, 1;
	setp.eq.b32 	%p61, %r710, 1;
	not.pred 	%p62, %p61;
	@%p62 bra 	$L__BB0_6;
	mov.b32 	%r737, 0;
	// begin inline asm
	sub.cc.u32 %r770, %r737, %r770;
	// end inline asm
	// begin inline asm
	subc.cc.u32 %r769, %r737, %r769;
	// end inline asm
	// begin inline asm
	subc.cc.u32 %r768, %r737, %r768;
	// end inline asm
	// begin inline asm
	subc.cc.u32 %r767, %r737, %r767;
	// end inline asm
	// begin inline asm
	subc.cc.u32 %r766, %r737, %r766;
	// end inline asm
	// begin inline asm
	subc.cc.u32 %r765, %r737, %r765;
	// end inline asm
	// begin inline asm
	subc.cc.u32 %r764, %r737, %r764;
	// end inline asm
	// begin inline asm
	subc.cc.u32 %r763, %r737, %r763;
	// end inline asm
	// begin inline asm
	subc.u32 %r735, %r737, %r737;
	// end inline asm
	mov.b32 	%r747, -662897337;
	mov.b32 	%r748, 1008765974;
	mov.b32 	%r749, 1752287885;
	mov.b32 	%r750, -1753126255;
	mov.b32 	%r751, -2122229667;
	mov.b32 	%r752, -1202698826;
	mov.b32 	%r753, -516841431;
	mov.b32 	%r754, 811880050;
	// begin inline asm
	{
	.reg .pred q;
	setp.ne.u32 q, %r735, 0;
	@q add.cc.u32 %r770, %r770, %r747;
	@q addc.cc.u32 %r769, %r769, %r748;
	@q addc.cc.u32 %r768, %r768, %r749;
	@q addc.cc.u32 %r767, %r767, %r750;
	@q addc.cc.u32 %r766, %r766, %r751;
	@q addc.cc.u32 %r765, %r765, %r752;
	@q addc.cc.u32 %r764, %r764, %r753;
	@q addc.u32 %r763, %r763, %r754;
	}
	// end inline asm
$L__BB0_6:
	ld.param.u64 	%rd104, [_Z7map_dbgRN5bn2545G1AffERKNS_2FpEPS2_Pi_param_0];
	cvta.to.local.u64 	%rd103, %rd104;
	st.local.v4.u32 	[%rd103], {%r58, %r59, %r60, %r61};
	st.local.v4.u32 	[%rd103+16], {%r74, %r75, %r76, %r77};
	st.local.v4.u32 	[%rd103+32], {%r770, %r769, %r768, %r767};
	st.local.v4.u32 	[%rd103+48], {%r766, %r765, %r764, %r763};
	ret;

}
